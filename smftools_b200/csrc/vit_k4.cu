// Viterbi kernels, K = 4
#define SMF_K 4
#include "vit_launch.inc"
