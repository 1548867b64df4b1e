// Viterbi kernels, K = 2
#define SMF_K 2
#include "vit_launch.inc"
