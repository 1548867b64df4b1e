// short chunk lengths (9, 11, 13, 15) of the K = 2 posterior kernel
#define SMF_K 2
#define SMF_FB2_EXTRA 2
#include "fb2_launch.inc"
