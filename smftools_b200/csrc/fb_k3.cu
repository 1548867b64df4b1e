// generic forward-backward kernels, K = 3
#define SMF_K 3
#include "fb_launch.inc"
