// fast-path forward-backward kernels, K = 3
#define SMF_K 3
#include "fb2_launch.inc"
