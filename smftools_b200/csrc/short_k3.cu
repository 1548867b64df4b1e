#define SMF_K 3
#include "short_launch.inc"
