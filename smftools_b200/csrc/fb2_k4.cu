// fast-path forward-backward kernels, K = 4
#define SMF_K 4
#include "fb2_launch.inc"
