#define SMF_K 2
#include "short_launch.inc"
