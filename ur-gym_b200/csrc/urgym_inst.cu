// urgym_inst.cu -- compiled once per (task, geometry): -DURGYM_INST_TASK=t -DURGYM_INST_GEOM=g (see Makefile).
#include "urgym_kernels.cuh"

#ifndef URGYM_INST_TASK
#error "define URGYM_INST_TASK and URGYM_INST_GEOM"
#endif
#define CAT_(a, b, c, d) a##b##_##d
#define NAME(prefix, T, G) CAT_(prefix, T, _, G)

cudaError_t NAME(urgym_inst_step_, URGYM_INST_TASK, URGYM_INST_GEOM)(const ModelConst &M, const StepArgs &A, cudaStream_t s) {
    return launch_step<URGYM_INST_TASK, URGYM_INST_GEOM>(M, A, s);
}
cudaError_t NAME(urgym_inst_reset_, URGYM_INST_TASK, URGYM_INST_GEOM)(const ModelConst &M, const AuxArgs &A, cudaStream_t s) {
    return launch_reset<URGYM_INST_TASK, URGYM_INST_GEOM>(M, A, s);
}
cudaError_t NAME(urgym_inst_autoreset_, URGYM_INST_TASK, URGYM_INST_GEOM)(const ModelConst &M, const AuxArgs &A, cudaStream_t s) {
    return launch_autoreset<URGYM_INST_TASK, URGYM_INST_GEOM>(M, A, s);
}
cudaError_t NAME(urgym_inst_refresh_, URGYM_INST_TASK, URGYM_INST_GEOM)(const ModelConst &M, const AuxArgs &A, cudaStream_t s) {
    return launch_refresh<URGYM_INST_TASK, URGYM_INST_GEOM>(M, A, s);
}
cudaError_t NAME(urgym_inst_prepare_, URGYM_INST_TASK, URGYM_INST_GEOM)(const ModelConst &M, const AuxArgs &A, cudaStream_t s) {
    return prepare_kernels<URGYM_INST_TASK, URGYM_INST_GEOM>(M, A, s);
}
// the kernels that only depend on the task are instantiated here too, once per geometry: every kernel a handle runs is
// then compiled with the same floating-point flags (Makefile: FASTMATH for the capsule geometry), so the episode cache
// rebuilt after a state injection is bit-identical to the one the handle's reset kernels write
cudaError_t NAME(urgym_inst_observe_, URGYM_INST_TASK, URGYM_INST_GEOM)(const ModelConst &M, const AuxArgs &A, cudaStream_t s) {
    return launch_observe<URGYM_INST_TASK, URGYM_INST_GEOM>(M, A, s);
}
cudaError_t NAME(urgym_inst_derive_, URGYM_INST_TASK, URGYM_INST_GEOM)(const ModelConst &M, const AuxArgs &A, cudaStream_t s) {
    return launch_derive<URGYM_INST_TASK, URGYM_INST_GEOM>(M, A, s);
}
