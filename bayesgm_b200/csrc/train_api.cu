// Translation unit of the EGM / iterative training entry points (train_api.cuh) and of the layered
// training engine (layered_api.cuh), which shares train.cuh's discriminator kernel.
#include <string>
#include "train_api.cuh"
#include "layered_api.cuh"
