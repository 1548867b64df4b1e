// generic forward-backward kernels, K = 4
#define SMF_K 4
#include "fb_launch.inc"
