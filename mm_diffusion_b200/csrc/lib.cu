// Single translation unit of libmmdiff.so (kernels are header-defined; one TU avoids duplicate definitions).
#include "ops.cu"
#include "model.cu"
