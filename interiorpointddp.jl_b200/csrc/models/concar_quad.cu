// Built-in workload 'concar_quad': instantiates the IPDDP2 kernels for the generated model struct.
#include "../models_gen/concar_quad.cuh"
#include "../model_register.cuh"
IPDDP_REGISTER_MODEL(Model_concar_quad, ipddp_vtable_concar_quad)
