// Built-in workload 'double_integrator': instantiates the IPDDP2 kernels for the generated model struct.
#include "../models_gen/double_integrator.cuh"
#include "../model_register.cuh"
IPDDP_REGISTER_MODEL(Model_double_integrator, ipddp_vtable_double_integrator)
