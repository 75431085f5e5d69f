// Built-in workload 'cartpole': instantiates the IPDDP2 kernels for the generated model struct.
#include "../models_gen/cartpole.cuh"
#include "../model_register.cuh"
IPDDP_REGISTER_MODEL(Model_cartpole, ipddp_vtable_cartpole)
