#define SMF_K 4
#include "short_launch.inc"
