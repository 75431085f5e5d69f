// Built-in workload 'acrobot': instantiates the IPDDP2 kernels for the generated model struct.
#include "../models_gen/acrobot.cuh"
#include "../model_register.cuh"
IPDDP_REGISTER_MODEL(Model_acrobot, ipddp_vtable_acrobot)
