// placeholder until the operator kernels land (next commit): symbols exist, calls fail loudly.
#include "../../include/ddnm_b200.h"
#include "api_util.cuh"
using namespace ddnm;
extern "C" {
#define NYI DDNM_API_BEGIN throw Error("operator kernels not built yet"); DDNM_API_END
int ddnm_operator_create(const ddnm_operator_desc*, void**) { NYI }
long long ddnm_operator_y_dim(void*) { return -1; }
int ddnm_operator_A(void*, const float*, int, float*, void*) { NYI }
int ddnm_operator_A_pinv(void*, const float*, int, float*, void*) { NYI }
int ddnm_operator_project(void*, const float*, const float*, int, float*, void*) { NYI }
int ddnm_operator_lambda(void*, const float*, int, float, float, float, float, float*, void*) { NYI }
int ddnm_operator_lambda_noise(void*, const float*, const float*, int, float, float, float, float, float*, void*) { NYI }
int ddnm_operator_destroy(void*) { NYI }
int ddnm_sample(void*, void*, const ddnm_schedule*, const float*, const float*, const float*, int, float*, float*, void*) { NYI }
}
