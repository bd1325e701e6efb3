/* tests/abi_main.c -- the C ABI from plain C (C99, no C++): the header must compile as C, every entry point must link, and the
 * calls that need no device must behave. With a device (argv[1] = "gpu") it also runs the batched loop of INTEGRATION.md section 3. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nuslam_b200.h"

/* taking the address of every entry point makes the linker resolve all of them */
typedef void (*entry_fn)(void);
static const entry_fn k_entry_points[] = {
    (entry_fn) nuslam_last_error, (entry_fn) nuslam_version, (entry_fn) nuslam_ekf_default_config, (entry_fn) nuslam_ekf_create,
    (entry_fn) nuslam_ekf_destroy, (entry_fn) nuslam_ekf_bind_state, (entry_fn) nuslam_ekf_device_pointers, (entry_fn) nuslam_ekf_init,
    (entry_fn) nuslam_ekf_set_state, (entry_fn) nuslam_ekf_get_state, (entry_fn) nuslam_ekf_predict, (entry_fn) nuslam_ekf_associate,
    (entry_fn) nuslam_ekf_initialize_landmark, (entry_fn) nuslam_ekf_update, (entry_fn) nuslam_ekf_measurement_model,
    (entry_fn) nuslam_ekf_step, (entry_fn) nuslam_ekf_step_async, (entry_fn) nuslam_ekf_wait_async, (entry_fn) nuslam_ekf_scan_step,
    (entry_fn) nuslam_ekf_map_to_odom, (entry_fn) nuslam_ekf_synchronize, (entry_fn) nuslam_cartesian2polar,
    (entry_fn) nuslam_normalize_angle, (entry_fn) nuslam_diffdrive_step, (entry_fn) nuslam_diffdrive_convert_twist,
    (entry_fn) nuslam_world_step, (entry_fn) nuslam_integrate_twist, (entry_fn) nuslam_scan_detect, (entry_fn) nuslam_classify_and_fit,
    (entry_fn) nuslam_ekf_get_stream, (entry_fn) nuslam_ekf_set_ids, (entry_fn) nuslam_ekf_step_async_packed, (entry_fn) nuslam_ekf_async_dry_run, (entry_fn) nuslam_ekf_error_stats, (entry_fn) nuslam_scan_set_fit, (entry_fn) nuslam_scan_last_fallbacks,
    (entry_fn) nuslam_tail_launch,
};

int main(int argc, char ** argv)
{
    nuslam_ekf_config cfg;
    nuslam_ekf * h = NULL;
    size_t k;
    for (k = 0; k < sizeof(k_entry_points) / sizeof(k_entry_points[0]); ++k)
        if (!k_entry_points[k]) return 1;
    nuslam_ekf_default_config(&cfg, 12);
    printf("VERSION %d\nENTRY_POINTS %d\n", nuslam_version(), (int) (sizeof(k_entry_points) / sizeof(k_entry_points[0])));
    printf("DEFAULT n=%d mode=%d Q00=%g R00=%g amin=%g amax=%g options=%u prior=%.1f\n", cfg.n_landmarks, cfg.mode, cfg.Q[0], cfg.R[0], cfg.assoc_min,
           cfg.assoc_max, cfg.options, cfg.landmark_prior);
    if (nuslam_ekf_create(NULL, 4, 0, NULL, &h) == NUSLAM_OK) return 2;   /* null config must be refused */
    printf("NULLCFG %s\n", nuslam_last_error());
    if (argc > 1 && strcmp(argv[1], "gpu") == 0)
    {
        enum { B = 4, N = 12, LEN = 3 + 2 * N, M = 2 };
        double robot[B * 3] = {0}, tw[B * 3], z[B * M * 2], x[B * LEN];
        int32_t ids[B * M], status[B], t, b;
        cfg.mode = NUSLAM_MODE_FAST;
        if (nuslam_ekf_create(&cfg, B, 0, NULL, &h) != NUSLAM_OK)
        {
            printf("CREATE FAILED %s\n", nuslam_last_error());
            return 3;
        }
        if (nuslam_ekf_init(h, robot, NULL, NUSLAM_HOST) != NUSLAM_OK) return 4;
        for (t = 0; t < 3; ++t)
        {
            for (b = 0; b < B; ++b)
            {
                tw[3 * b] = 0.02;
                tw[3 * b + 1] = 0.007;
                tw[3 * b + 2] = 0.0;
                z[(b * M + 0) * 2] = 0.5;
                z[(b * M + 0) * 2 + 1] = 0.3;
                z[(b * M + 1) * 2] = 0.7;
                z[(b * M + 1) * 2 + 1] = -0.4;
                ids[b * M] = 1;
                ids[b * M + 1] = 2;
            }
            if (nuslam_ekf_step(h, tw, z, ids, M, NULL, NUSLAM_HOST) != NUSLAM_OK)
            {
                printf("STEP FAILED %s\n", nuslam_last_error());
                return 5;
            }
        }
        if (nuslam_ekf_get_state(h, x, NULL, NULL, status, NUSLAM_HOST) != NUSLAM_OK) return 6;
        printf("STATE %.12f %.12f %.12f status %d\n", x[0], x[1], x[2], status[0]);
        nuslam_ekf_destroy(h);
    }
    return 0;
}
