// extra chunk lengths (13, 21) of the K = 3 posterior kernel
#define SMF_K 3
#define SMF_FB2_EXTRA 1
#include "fb2_launch.inc"
