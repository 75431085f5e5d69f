// Built-in stage chain 'ragged' (state and control sizes change along the horizon): instantiates the IPDDP2 kernels
// for the generated composite model.
#include "../models_gen/ragged.cuh"
#include "../model_register.cuh"
IPDDP_REGISTER_MODEL(Model_ragged, ipddp_vtable_ragged)
