/*
 * main.c -- `qsim`, the C host driver.  Drop-in for the reference CLIs:
 *   C program   : <exe> <circuit_file> <number_of_measurement>   (quantum_simulator.c:39-43)
 *   CUDA programs: <exe> <circuit_file>                           (quantum_simulator_naive.cu:135-139)
 * stdout line 1 is the elapsed seconds "%lf\n" exactly as the reference prints
 * it (quantum_simulator.c:248).  Everything else is opt-in and follows the
 * formats the reference left commented out:
 *   --dump-amplitudes   "%llu : %f + %f i" per non-zero amplitude, then
 *                       "MOST LIKELY MEASUREMENT: %llu (%f)"   (naive.cu:207-216)
 *   <number_of_measurement> > 0 with --shots: "MEASUREMENT: <bits> (%ld)" (quantum_simulator.c:68-73)
 *   --save-state FILE / --load-state FILE   raw shard + qubit map after / before the circuit (checkpoint;
 *                       the reference keeps the state only in memory, quantum_simulator.c:75)
 *   --plan-only [--gpus N]   host-side fusion only (no GPU needed): one JSON line with the passes, rounds, ops and
 *                       qubit exchanges the circuit would run as on N ranks, instead of running it
 * Errors go to stdout followed by exit(1), like the reference (:56,:129,:213-219).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <time.h>

#include "qsim_b200.h"

static double now_s(void)
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + tv.tv_usec * 1e-6;
}

static void putb(unsigned long long n, int len) /* MSB first, quantum_simulator.c:285-293 */
{
    for (int k = len - 1; k >= 0; k--) putchar('0' + (int)((n >> k) & 1ULL));
}

static void usage(const char *exe)
{
    printf("QUANTUM CIRCUIT SIMULATOR\n");
    printf("Usage: %s <circuit_file_name> <number_of_measurement>\n", exe);
}

static void format_help(void)
{
    printf("Input format: \n\n");
    printf("OPENQASM 3.0;\n");
    printf("include \"stdgates.inc\";\n");
    printf("qubit[<num_qubit>] q; or qubit q[<num_qubit>]; \\\\single quantum register \n");
    printf("<quantum_circuit>\n\n");
    printf("Supported operations: cx, x, sx, z, s, sdg, t, tdg, rz, h\n");
    printf("(extensions: y p rx ry u cz cy ch cp swap ccx, gate definitions, ctrl/negctrl/inv/pow modifiers, gphase, pi expressions)\n");
}

int main(int argc, char **argv)
{
    const char *file = NULL, *dump_bin = NULL, *save_state = NULL, *load_state = NULL;
    long num_m = 0;
    int precision = 32, dump = 0, shots = 0, prec_out = 0, profile = 0, sweep = 0, have_m = 0, plan_only = 0, gpus = 1;
    unsigned long long seed = 0; int have_seed = 0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--precision") && i + 1 < argc) precision = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--dump-amplitudes")) dump = 1;
        else if (!strcmp(argv[i], "--precision-out") && i + 1 < argc) prec_out = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--dump-bin") && i + 1 < argc) dump_bin = argv[++i];
        else if (!strcmp(argv[i], "--save-state") && i + 1 < argc) save_state = argv[++i];
        else if (!strcmp(argv[i], "--load-state") && i + 1 < argc) load_state = argv[++i];
        else if (!strcmp(argv[i], "--shots")) shots = 1;
        else if (!strcmp(argv[i], "--seed") && i + 1 < argc) { seed = strtoull(argv[++i], NULL, 10); have_seed = 1; }
        else if (!strcmp(argv[i], "--profile")) profile = 1;
        else if (!strcmp(argv[i], "--plan-only")) plan_only = 1;
        else if (!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--sweep")) sweep = 1;
        else if (!file) file = argv[i];
        else if (!have_m) { num_m = atol(argv[i]); have_m = 1; }
    }
    if (!file) { usage(argv[0]); return 1; }
    if (!have_seed) seed = (unsigned long long)time(NULL);   /* srand(time(NULL)), :46 */

    int nq = 0; qsb_gate_t *gates = NULL; size_t n = 0;
    int rc = qsb_parse_qasm_file(file, &nq, &gates, &n);
    if (rc == QSB_ERR_IO) { printf("ERROR: cannot open circuit file\n"); return 1; }
    if (rc) {
        printf("%s\n", qsb_last_error());
        if (rc == QSB_ERR_PARSE) format_help();
        printf("ERROR while parsing quantum circuit\n");
        return 1;
    }

    double t0 = now_s();
    qsb_options_t o; qsb_options_default(&o);
    o.precision = precision == 64 ? QSB_F64 : QSB_F32;
    if (sweep) o.mode = QSB_MODE_SWEEP;
    if (plan_only) {
        qsb_run_stats_t st;
        o.world_size = gpus;
        if (qsb_plan_dry_run(nq, &o, gates, n, &st)) { printf("%s\n", qsb_last_error()); return 1; }
        printf("{\"qubits\": %d, \"gates\": %llu, \"ranks\": %d, \"precision\": %d, \"passes\": %u, \"rounds\": %u, \"device_ops\": %llu, "
               "\"exchanges\": %u, \"bytes_moved_per_rank\": %llu, \"bytes_exchanged_per_rank\": %llu, \"plan_ms\": %.3f, \"lib\": \"%s\"}\n",
               nq, (unsigned long long)st.source_gates, gpus, precision == 64 ? 64 : 32, st.passes, st.rounds, (unsigned long long)st.device_ops,
               st.swaps, (unsigned long long)st.bytes_moved, (unsigned long long)st.bytes_exchanged, st.plan_ms, qsb_version());
        qsb_free(gates);
        return 0;
    }
    qsb_t *s = NULL;
    rc = qsb_create(&s, nq, &o);
    if (!rc && load_state) rc = qsb_load_state(s, load_state);
    if (!rc) rc = qsb_apply_gates(s, gates, n);
    if (!rc && save_state) rc = qsb_save_state(s, save_state);
    if (rc) { printf("%s\n", qsb_last_error()); return 1; }
    printf("%lf\n", now_s() - t0);

    if (profile) {
        qsb_run_stats_t st; qsb_last_run_stats(s, &st);
        printf("{\"qubits\": %d, \"gates\": %llu, \"passes\": %u, \"rounds\": %u, \"device_ms\": %.6f, \"plan_ms\": %.3f, "
               "\"bytes_moved\": %llu, \"gbps\": %.1f, \"gates_per_sec\": %.1f}\n",
               nq, (unsigned long long)st.source_gates, st.passes, st.rounds, st.device_ms, st.plan_ms,
               (unsigned long long)st.bytes_moved, st.device_ms > 0 ? st.bytes_moved / st.device_ms * 1e-6 : 0.0,
               st.device_ms > 0 ? st.source_gates / st.device_ms * 1e3 : 0.0);
    }
    const unsigned long long N = 1ULL << nq;
    if (dump || dump_bin) {
        const unsigned long long chunk = 1ULL << 20;
        double *buf = (double *)malloc(sizeof(double) * 2 * (N < chunk ? N : chunk));
        FILE *fb = dump_bin ? fopen(dump_bin, "wb") : NULL;
        if (!buf || (dump_bin && !fb)) { printf("Malloc error\n"); return 1; }
        for (unsigned long long f = 0; f < N; f += chunk) {
            unsigned long long c = N - f < chunk ? N - f : chunk;
            if (qsb_download(s, buf, f, c)) { printf("%s\n", qsb_last_error()); return 1; }
            if (fb) fwrite(buf, sizeof(double), 2 * c, fb);
            if (dump) for (unsigned long long k = 0; k < c; k++) {
                double re = buf[2 * k], im = buf[2 * k + 1];
                if (re * re + im * im > 0.0) {
                    if (prec_out > 0) printf("%llu : %.*g + %.*g i\n", f + k, prec_out, re, prec_out, im);
                    else printf("%llu : %f + %f i\n", f + k, re, im);
                }
            }
        }
        if (fb) fclose(fb);
        free(buf);
        if (dump) {
            double norm, p; uint64_t idx;
            if (qsb_norm_argmax(s, &norm, &idx, &p)) { printf("%s\n", qsb_last_error()); return 1; }
            printf("MOST LIKELY MEASUREMENT: %llu (%f)\n", (unsigned long long)idx, p);
        }
    }
    if (shots && num_m > 0) {
        uint64_t *out = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)num_m);
        if (!out) { printf("Malloc error\n"); return 1; }
        if (qsb_sample(s, seed, (int)num_m, out)) { printf("%s\n", qsb_last_error()); return 1; }
        for (long i = 0; i < num_m; i++) {
            printf("MEASUREMENT: "); putb(out[i], nq); printf(" (%ld)\n", (long)out[i]);
        }
        free(out);
    }
    qsb_destroy(s);
    qsb_free(gates);
    return 0;
}
