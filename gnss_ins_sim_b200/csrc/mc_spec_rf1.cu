#define B2_RF 1
#define B2_SPEC_NAME launch_mc_spec_rf1
#include "mc_spec_launch.cuh"
