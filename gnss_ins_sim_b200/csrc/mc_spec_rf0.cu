#define B2_RF 0
#define B2_SPEC_NAME launch_mc_spec_rf0
#include "mc_spec_launch.cuh"
