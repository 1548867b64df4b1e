// sequential (thread-per-read) E-step kernels, K = 3
#define SMF_K 3
#include "seq_launch.inc"
