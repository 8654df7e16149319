// imx_kernels_inst.cu — ahead-of-time instantiations of the step / rollout kernels for ONE (tile width, network family)
// pair, selected by -DIMX_INST_MPAD=<2|4|8|16|32> -DIMX_INST_DIV=<0|1>.  The build compiles the nine pairs as separate
// objects in parallel (the instantiations are most of the library's compile time) and links them with imx_api.cu.
#include "imx_kernels.cuh"

#ifndef IMX_INST_MPAD
#error "compile with -DIMX_INST_MPAD=<tile width> -DIMX_INST_DIV=<0|1>"
#endif

using namespace imx;

#define IMX_CAT2(a, b, c) imx_pick_kernels_##a##_##b
#define IMX_CAT(a, b) IMX_CAT2(a, b, )
#define IMX_PICK_NAME IMX_CAT(IMX_INST_MPAD, IMX_INST_DIV)

// small: lead times <= 4 and at most one history slot (every shipped preset); few: at most two children per node
void IMX_PICK_NAME(bool small, bool few, KernelSet* ks) {
    constexpr int M_PAD = IMX_INST_MPAD;
#if IMX_INST_DIV
    if (small && few) { ks->tma_fn = step_kernel_tma<M_PAD, 4, 1, 2, true>; ks->tma_many_fn = step_kernel_tma_many<M_PAD, 4, 1, 2, true>; }
    else if (small)   { ks->tma_fn = step_kernel_tma<M_PAD, 4, 1, 8, true>; ks->tma_many_fn = step_kernel_tma_many<M_PAD, 4, 1, 8, true>; }
    else if (few)     { ks->tma_fn = step_kernel_tma<M_PAD, 8, 8, 2, true>; ks->tma_many_fn = step_kernel_tma_many<M_PAD, 8, 8, 2, true>; }
    else              { ks->tma_fn = step_kernel_tma<M_PAD, 8, 8, 8, true>; ks->tma_many_fn = step_kernel_tma_many<M_PAD, 8, 8, 8, true>; }
    if (small && few) { ks->step_fn = step_kernel<M_PAD, 4, 1, 2, true>; ks->rollout_fn = rollout_kernel<M_PAD, 4, 2, true>; }
    else if (small)   { ks->step_fn = step_kernel<M_PAD, 4, 1, 8, true>; ks->rollout_fn = rollout_kernel<M_PAD, 4, 8, true>; }
    else if (few)     { ks->step_fn = step_kernel<M_PAD, 8, 8, 2, true>; ks->rollout_fn = rollout_kernel<M_PAD, 8, 2, true>; }
    else              { ks->step_fn = step_kernel<M_PAD, 8, 8, 8, true>; ks->rollout_fn = rollout_kernel<M_PAD, 8, 8, true>; }
#else
    (void)few;
    ks->tma_fn = small ? step_kernel_tma<M_PAD, 4, 1, 1, false> : step_kernel_tma<M_PAD, 8, 8, 1, false>;
    ks->tma_many_fn = small ? step_kernel_tma_many<M_PAD, 4, 1, 1, false> : step_kernel_tma_many<M_PAD, 8, 8, 1, false>;
    if (small) { ks->step_fn = step_kernel<M_PAD, 4, 1, 1, false>; ks->rollout_fn = rollout_kernel<M_PAD, 4, 1, false>; }
    else       { ks->step_fn = step_kernel<M_PAD, 8, 8, 1, false>; ks->rollout_fn = rollout_kernel<M_PAD, 8, 1, false>; }
#endif
}
