// imx_step_pipe.cuh — placeholder (filled in below in this round).
#pragma once

#include "imx_step_tma.cuh"
