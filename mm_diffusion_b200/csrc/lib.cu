// Single translation unit of libmmdiff.so (kernels are header-defined; one TU avoids duplicate definitions).
#include "ops.cu"
#include "ops_bwd.cu"
#include "model.cu"
