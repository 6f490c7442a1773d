/*
 * gpu_lane_reloc_check.c -- TEST TOOL (B200 box, no Python): hardware check and A/B of the end-of-pass lane
 * relocation (tiled_plan.cpp, build_rounds) through the C ABI.
 *   1. parity: a circuit file is run by the CPU oracle (oracle/_build/liboracle.so, oc_run_file) and on the GPU
 *      in f32 and f64, with the default planner, with the planner before the lane relocation and the CX -> controlled-phase
 *      rewrite (reserved[6] = 3, reserved[4] = 5) and with reserved[6] = 4 (relocation of the qubits in conflict only);
 *      prints max |amplitude difference| per run and exits 1 above 1e-5 (f32) / 1e-12 (f64).
 *   2. timing: every further file is planned once per planner setting (old / lane relocation only / default) and executed
 *      `reps` times from |0...0>; prints the
 *      CUDA-event time per execution (minimum and all), passes and rounds.
 * Usage: gpu_lane_reloc_check <parity.qasm> <reps> <bench.qasm[:64]>...
 * Build: tests/tools/Makefile.  The oracle is the checker here, never the thing measured.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qsim_b200.h"

int oc_run_file(const char *path, double **state, int *num_q_out);   /* oracle/qsim_oracle.c */
void oc_free(void *p);

static int check(int rc, const char *what)
{
    if (rc) { printf("FAIL %s: %s\n", what, qsb_last_error()); exit(2); }
    return rc;
}

int main(int argc, char **argv)
{
    if (argc < 3) { printf("usage: %s <parity.qasm> <reps> <bench.qasm[:64]>...\n", argv[0]); return 2; }
    int bad = 0;
    /* ---- 1. parity against the oracle ---- */
    {
        int nq = 0; qsb_gate_t *gates = NULL; size_t n = 0;
        check(qsb_parse_qasm_file(argv[1], &nq, &gates, &n), "parse");
        double *want = NULL; int nq2 = 0;
        if (oc_run_file(argv[1], &want, &nq2) || nq2 != nq) { printf("FAIL oracle on %s\n", argv[1]); return 2; }
        const uint64_t N = 1ULL << nq;
        double *got = (double *)malloc(sizeof(double) * 2 * N);
        static const int pol[] = {0, 3, 4}, cxh[] = {0, 5, 0};
        for (int prec = 32; prec <= 64; prec += 32) for (int k = 0; k < 3; k++) {
            qsb_options_t o; qsb_options_default(&o);
            o.precision = prec == 64 ? QSB_F64 : QSB_F32; o.reserved[6] = pol[k]; o.reserved[4] = cxh[k];
            qsb_t *s = NULL;
            check(qsb_create(&s, nq, &o), "create");
            check(qsb_apply_gates(s, gates, n), "apply");
            qsb_run_stats_t st; qsb_last_run_stats(s, &st);
            check(qsb_download(s, got, 0, N), "download");
            double err = 0;
            for (uint64_t i = 0; i < 2 * N; i++) { const double d = fabs(got[i] - want[i]); if (d > err) err = d; }
            const double tol = prec == 64 ? 1e-12 : 1e-5;
            printf("{\"parity\": \"%s\", \"qubits\": %d, \"gates\": %zu, \"precision\": %d, \"lane_policy\": %d, \"cx_kept\": %d, \"passes\": %u, \"rounds\": %u, "
                   "\"max_abs_err\": %.3e, \"tol\": %.0e, \"ok\": %s}\n", argv[1], nq, n, prec, pol[k], cxh[k] == 5, st.passes, st.rounds, err, tol, err <= tol ? "true" : "false");
            fflush(stdout);
            if (!(err <= tol)) bad = 1;
            qsb_destroy(s);
        }
        free(got); oc_free(want); qsb_free(gates);
    }
    /* ---- 2. A/B timing ---- */
    const int reps = atoi(argv[2]);
    for (int a = 3; a < argc; a++) {
        char path[1024]; snprintf(path, sizeof path, "%s", argv[a]);
        int prec = 32;
        char *colon = strrchr(path, ':');
        if (colon) { prec = atoi(colon + 1); *colon = 0; }
        int nq = 0; qsb_gate_t *gates = NULL; size_t n = 0;
        check(qsb_parse_qasm_file(path, &nq, &gates, &n), "parse");
        static const int pol[] = {3, 0, 0}, cxh[] = {5, 5, 0};   /* the planner before both changes, lane relocation only, the default */
        for (int k = 0; k < 3; k++) {
            qsb_options_t o; qsb_options_default(&o);
            o.precision = prec == 64 ? QSB_F64 : QSB_F32; o.reserved[6] = pol[k]; o.reserved[4] = cxh[k];
            qsb_t *s = NULL;
            check(qsb_create(&s, nq, &o), "create");
            qsb_plan_t *p = NULL;
            check(qsb_plan_create(s, gates, n, &p), "plan");
            double best = 1e30; char all[512]; all[0] = 0;
            qsb_run_stats_t st; memset(&st, 0, sizeof st);
            for (int r = 0; r < reps; r++) {
                check(qsb_reset(s), "reset");
                check(qsb_execute(s, p), "execute");
                qsb_last_run_stats(s, &st);
                if (r > 0 && st.device_ms < best) best = st.device_ms;     /* the first execution warms up */
                snprintf(all + strlen(all), sizeof all - strlen(all), "%s%.3f", r ? ", " : "", st.device_ms);
            }
            double norm = 0, pmax = 0; uint64_t idx = 0;
            check(qsb_norm_argmax(s, &norm, &idx, &pmax), "norm");
            printf("{\"bench\": \"%s\", \"qubits\": %d, \"gates\": %zu, \"precision\": %d, \"lane_policy\": %d, \"cx_kept\": %d, \"passes\": %u, \"rounds\": %u, "
                   "\"ms_min\": %.3f, \"ms_all\": [%s], \"norm\": %.9f}\n", path, nq, n, prec, pol[k], cxh[k] == 5, st.passes, st.rounds, best, all, norm);
            fflush(stdout);
            if (fabs(norm - 1.0) > (prec == 64 ? 1e-9 : 1e-3)) bad = 1;
            qsb_plan_destroy(p);
            qsb_destroy(s);
        }
        qsb_free(gates);
    }
    return bad;
}
