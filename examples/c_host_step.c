/*
 * A plain C host driving the env step through the C ABI (include/marllb_b200.h) -- no CUDA headers, no Python,
 * no torch: host buffers in, host buffers out.  This is the call sequence a maintainer's binding would make in
 * place of constructing E LoadBalanceEnv objects and looping env.step()
 * (simulation-mode/problem-03-rl-environment/examples/random_policy.py:52-61).
 *
 *   gcc -O2 -Iinclude examples/c_host_step.c -o /tmp/c_host_step marllb_b200/libmarllb_b200.so -Wl,-rpath,$PWD/marllb_b200 -lm
 *   /tmp/c_host_step [envs] [servers] [steps]
 *
 * Prints per-step mean reward and, at the end, a checksum line "checksum <sum n_flow_on> <sum reward>" that
 * tests/test_gpu_c_host.py compares with the same run made through the Python classes.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "marllb_b200.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc__ = (call);                                                               \
        if (rc__ != MLB_OK) {                                                            \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc__, mlb_last_error(env));          \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

int main(int argc, char **argv) {
    const int E = argc > 1 ? atoi(argv[1]) : 256, S = argc > 2 ? atoi(argv[2]) : 16, steps = argc > 3 ? atoi(argv[3]) : 40;
    mlb_env *env = NULL;
    mlb_config cfg;
    if (mlb_abi_version() != MLB_ABI_VERSION) {
        fprintf(stderr, "ABI mismatch\n");
        return 1;
    }
    mlb_config_default(&cfg);                    /* env.py:71-87 defaults: discrete weights {1, 1.5, 2}, jain on column 10 */
    cfg.num_envs = E;
    cfg.num_agents = 1;
    cfg.servers_per_agent = S;
    cfg.max_steps = steps;
    cfg.policy = MLB_POLICY_SED;
    if (mlb_create(&cfg, &env) != MLB_OK) {
        fprintf(stderr, "mlb_create: %s\n", mlb_last_error(NULL));
        return 1;
    }
    float *speeds = malloc(sizeof(float) * S);
    for (int j = 0; j < S; j++) speeds[j] = (j % 2 == 0) ? 1.0f : 2.0f;
    CHECK(mlb_set_speeds(env, speeds, S, MLB_HOST, NULL));
    /* synthetic Poisson arrivals on the device (training_pipeline.py:141-155 semantics), rho = 0.8 */
    const double rate = 2.0 * S, mean_work = 0.8 * 1.5 * S / rate;
    CHECK(mlb_gen_poisson(env, rate, mean_work, steps * 0.25 + 1.0, 1234, NULL));
    CHECK(mlb_reset(env, NULL, NULL));

    int32_t *action = malloc(sizeof(int32_t) * (size_t)E * S);
    float *obs = malloc(sizeof(float) * (size_t)E * S * 11);
    double *reward = malloc(sizeof(double) * E);
    uint8_t *done = malloc(E);
    double reward_total = 0.0;
    uint32_t lcg = 12345u;                       /* any policy: here a fixed pseudo-random one */
    for (int k = 0; k < steps; k++) {
        for (size_t i = 0; i < (size_t)E * S; i++) {
            lcg = lcg * 1664525u + 1013904223u;
            action[i] = (int32_t)((lcg >> 16) % 3u);
        }
        CHECK(mlb_step(env, action, MLB_HOST, obs, reward, done, MLB_HOST, NULL));
        CHECK(mlb_status(env, NULL));            /* also synchronises the (default) stream: outputs are valid */
        double mean = 0.0;
        for (int e = 0; e < E; e++) mean += reward[e];
        reward_total += mean;
        if (k % 10 == 0 || k == steps - 1)
            printf("step %3d  mean reward %.6f  done %d  obs[0][0] = [%g %g %g ...]\n", k + 1, mean / E, (int)done[0],
                   obs[0], obs[1], obs[2]);
    }
    int32_t *n_on = malloc(sizeof(int32_t) * (size_t)E * S);
    CHECK(mlb_get_state(env, MLB_F_N_FLOW_ON, n_on, sizeof(int32_t) * (size_t)E * S, MLB_HOST));
    long long flows = 0;
    for (size_t i = 0; i < (size_t)E * S; i++) flows += n_on[i];
    printf("checksum %lld %.9f\n", flows, reward_total);
    printf("kernels launched: %lld\n", (long long)mlb_launch_count(env));
    mlb_destroy(env);
    free(speeds); free(action); free(obs); free(reward); free(done); free(n_on);
    return 0;
}
