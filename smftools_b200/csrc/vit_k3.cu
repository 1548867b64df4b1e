// Viterbi kernels, K = 3
#define SMF_K 3
#include "vit_launch.inc"
