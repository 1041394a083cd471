// Single translation unit for libtnqs_b200.so (kernels live in headers; one TU keeps one copy).
#include "engine.cu"
#include "capi.cu"
