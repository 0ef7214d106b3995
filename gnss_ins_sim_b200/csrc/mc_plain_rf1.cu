#define B2_RF 1
#define B2_PLAIN_NAME launch_mc_plain_rf1
#include "mc_plain_launch.cuh"
