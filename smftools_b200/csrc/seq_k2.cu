// sequential (thread-per-read) E-step kernels, K = 2
#define SMF_K 2
#include "seq_launch.inc"
