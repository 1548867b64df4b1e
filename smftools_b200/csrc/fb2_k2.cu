// fast-path forward-backward kernels, K = 2
#define SMF_K 2
#include "fb2_launch.inc"
