#define SMF_K 3
#include "fb2b_launch.inc"
