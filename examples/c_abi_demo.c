/* c_abi_demo.c — the C ABI (include/imx_b200.h) driven from plain C with nothing but the CUDA runtime: no Python, no
 * torch.  Runs one episode of the 4-stage multi-agent chain (inv_management.py:36-56, preset MA_6 of hyperparams.py) on
 * N environments with a replayed demand trace and a stored action trace, first with one imx_step call per period and
 * then again through imx_step_many, and writes the rewards and the last observation to a file.
 *
 *   c_abi_demo <in.bin> <out.bin>
 *   in.bin : int64 N, int64 T, then demand int32[N][T], then actions float64[T][N][4]
 *   out.bin: rewards float64[T][N][4], last observation float64[N][4][O], the same two blocks again from imx_step_many
 *
 * tests/test_gpu_c_example.py builds the inputs, runs this program and checks the outputs against the oracle.
 * Build (see __graft_entry__.build):  gcc -std=c99 examples/c_abi_demo.c -Iinclude -I$CUDA/include -Lmarl_for_im_b200
 *                                     -limx_b200 -L$CUDA/lib64 -lcudart -o examples/c_abi_demo */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "imx_b200.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define CHECK_IMX(x) do { if ((x) < 0) { fprintf(stderr, "%s: %s\n", #x, imx_last_error()); return 3; } } while (0)

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 1; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    long long N = 0, T = 0;
    if (fread(&N, 8, 1, f) != 1 || fread(&T, 8, 1, f) != 1) return 1;
    const int m = 4;
    int32_t* demand = (int32_t*)malloc((size_t)N * T * sizeof(int32_t));
    double* actions = (double*)malloc((size_t)T * N * m * sizeof(double));
    if (fread(demand, sizeof(int32_t), (size_t)N * T, f) != (size_t)(N * T)) return 1;
    if (fread(actions, sizeof(double), (size_t)T * N * m, f) != (size_t)(T * N * m)) return 1;
    fclose(f);

    /* Env(config): the host side of the reference constructor (MAIM_env.py:8-173) — explicit vectors, defaults applied */
    imx_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    if (imx_config_size() != (int)sizeof(cfg) || imx_abi_version() != IMX_ABI_VERSION) { fprintf(stderr, "header / library mismatch\n"); return 1; }
    cfg.kind = IMX_KIND_MAIM; cfg.num_nodes = m; cfg.num_periods = (int32_t)T; cfg.prev_length = 1;
    cfg.time_dependency = 1; cfg.prev_demand = 1; cfg.prev_actions = 0;
    cfg.standardise_state = 1; cfg.standardise_actions = 1; cfg.independent = 0;
    cfg.demand_dist = IMX_DIST_REPLAY_ONLY; cfg.device = 0; cfg.a = -1.0; cfg.b = 1.0; cfg.mu = 5.0; cfg.seed = 52;
    cfg.num_envs = N; cfg.env_offset = 0;
    {
        const double price[5] = {5, 4, 3, 2, 1}, stock[4] = {0.35, 0.3, 0.4, 0.2}, backlog[4] = {0.5, 0.7, 0.6, 0.9};
        const int delay[4] = {1, 2, 3, 1};
        int i;
        for (i = 0; i < m; ++i) {
            cfg.inv_init[i] = 10; cfg.inv_max[i] = 30; cfg.order_max[i] = 30;       /* order_max[i] = inv_max[i+1], last = its own (MAIM_env.py:58-62) */
            cfg.delay[i] = delay[i]; cfg.inv_target[i] = 0.0; cfg.stock_cost[i] = stock[i]; cfg.backlog_cost[i] = backlog[i];
        }
        for (i = 0; i <= m; ++i) cfg.price[i] = price[i];
    }
    imx_env* env = NULL;
    CHECK_IMX(imx_create(&cfg, &env));
    const int O = imx_obs_len(env);
    const size_t cells = (size_t)N * m;

    int32_t* d_demand; double *d_actions, *d_obs, *d_rew;
    CHECK_CUDA(cudaMalloc((void**)&d_demand, (size_t)N * T * sizeof(int32_t)));
    CHECK_CUDA(cudaMalloc((void**)&d_actions, (size_t)T * cells * sizeof(double)));
    CHECK_CUDA(cudaMalloc((void**)&d_obs, (size_t)T * cells * O * sizeof(double)));
    CHECK_CUDA(cudaMalloc((void**)&d_rew, (size_t)T * cells * sizeof(double)));
    CHECK_CUDA(cudaMemcpy(d_demand, demand, (size_t)N * T * sizeof(int32_t), cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(d_actions, actions, (size_t)T * cells * sizeof(double), cudaMemcpyHostToDevice));
    cudaStream_t stream;
    CHECK_CUDA(cudaStreamCreate(&stream));

    double* rew = (double*)malloc((size_t)T * cells * sizeof(double));
    double* obs_last = (double*)malloc(cells * O * sizeof(double));
    FILE* out = fopen(argv[2], "wb");
    if (!out) { perror(argv[2]); return 1; }
    int pass;
    for (pass = 0; pass < 2; ++pass) {
        CHECK_CUDA(cudaMemsetAsync(d_rew, 0, (size_t)T * cells * sizeof(double), stream));
        CHECK_IMX(imx_reset(env, d_demand, NULL, 0, 1, NULL, stream));           /* reset(customer_demand=...) */
        if (pass == 0) {
            long long t;
            for (t = 0; t < T; ++t)                                               /* obs, reward, done, info = env.step(action) */
                CHECK_IMX(imx_step(env, d_actions + t * cells, d_obs + t * cells * O, d_rew + t * cells, NULL, stream));
        } else {
            CHECK_IMX(imx_step_many(env, d_actions, (int)T, d_obs, d_rew, NULL, stream));   /* the same episode as one call */
        }
        if (imx_period(env) != T) { fprintf(stderr, "period %d after the episode\n", imx_period(env)); return 4; }
        CHECK_CUDA(cudaMemcpyAsync(rew, d_rew, (size_t)T * cells * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CHECK_CUDA(cudaMemcpyAsync(obs_last, d_obs + (size_t)(T - 1) * cells * O, cells * O * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CHECK_CUDA(cudaStreamSynchronize(stream));
        fwrite(rew, sizeof(double), (size_t)T * cells, out);
        fwrite(obs_last, sizeof(double), cells * O, out);
    }
    fclose(out);
    if (imx_step(env, d_actions, d_obs, d_rew, NULL, stream) != -6) { fprintf(stderr, "stepping past the end must return -6\n"); return 5; }
    printf("ok N=%lld T=%lld O=%d kernel_variant=%d launches=%lld\n", N, T, O, imx_kernel_variant(env), (long long)imx_launch_count());
    CHECK_IMX(imx_destroy(env));
    return 0;
}
