// Translation unit of the EGM / iterative training entry points (train_api.cuh).
#include <string>
#include "train_api.cuh"
