// sequential (thread-per-read) E-step kernels, K = 4
#define SMF_K 4
#include "seq_launch.inc"
