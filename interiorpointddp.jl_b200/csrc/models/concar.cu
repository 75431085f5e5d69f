// Built-in workload 'concar': instantiates the IPDDP2 kernels for the generated model struct.
#include "../models_gen/concar.cuh"
#include "../model_register.cuh"
IPDDP_REGISTER_MODEL(Model_concar, ipddp_vtable_concar)
