#define SMF_K 4
#include "fb2w_launch.inc"
