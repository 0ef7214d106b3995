#define B2_RF 0
#define B2_PLAIN_NAME launch_mc_plain_rf0
#include "mc_plain_launch.cuh"
