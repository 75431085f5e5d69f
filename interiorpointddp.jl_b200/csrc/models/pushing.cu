// Built-in workload 'pushing': instantiates the IPDDP2 kernels for the generated model struct.
#include "../models_gen/pushing.cuh"
#include "../model_register.cuh"
IPDDP_REGISTER_MODEL(Model_pushing, ipddp_vtable_pushing)
