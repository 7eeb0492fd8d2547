// placeholder until the network plan lands (next commit): keeps every symbol of include/mmdiff.h exported.
#include "host.cuh"
using namespace mmd;
extern "C" {
int mmd_model_create(const MmdConfig*, MmdModel**) { return fail(MMD_ESTATE, "model plan not built yet"); }
int mmd_model_destroy(MmdModel*) { return MMD_OK; }
int mmd_model_num_params(const MmdModel*) { return 0; }
int mmd_model_param_info(const MmdModel*, int, const char**, int*, int64_t*) { return fail(MMD_ESTATE, "n/a"); }
int mmd_model_set_param(MmdModel*, const char*, const float*, int64_t, void*) { return fail(MMD_ESTATE, "n/a"); }
int mmd_model_num_shifts(const MmdModel*) { return 0; }
int mmd_model_shift_bound(const MmdModel*, int) { return 0; }
size_t mmd_model_workspace_bytes(const MmdModel*, int) { return 0; }
int mmd_model_num_launches(const MmdModel*, int) { return 0; }
int mmd_model_forward(MmdModel*, int, const float*, const float*, const float*, const int32_t*, float*, float*, void*) { return fail(MMD_ESTATE, "n/a"); }
}
