// Translation unit of the BGM / HMC entry points (hmc_api.cuh).
#include <string>
#include "hmc_api.cuh"
