// generic forward-backward kernels, K = 2
#define SMF_K 2
#include "fb_launch.inc"
