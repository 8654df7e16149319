// imx_kernels.cuh — the set of ahead-of-time kernel instantiations one env handle dispatches to.
#pragma once

#include "imx_rollout.cuh"
#include "imx_step_tma.cuh"

namespace imx {

typedef void (*step_fn_t)(const StepArgs);
typedef void (*tma_fn_t)(const StepArgs, const TileLayout);
typedef void (*rollout_fn_t)(const StepArgs, const RolloutArgs);

struct KernelSet {
    step_fn_t step_fn = nullptr;          // direct kernel (tails, unaligned buffers, diagnostics)
    tma_fn_t tma_fn = nullptr;            // TMA-staged tiles, one period per launch
    tma_fn_t tma_many_fn = nullptr;       // TMA-staged tiles, K periods per launch (imx_step_many)
    rollout_fn_t rollout_fn = nullptr;    // fused base-stock episode
};

}  // namespace imx
