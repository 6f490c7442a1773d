/*
 * qsim_oracle.c -- CPU restatement of the reference state-vector path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under gpu_quantum_simulator_b200/ may
 * include, link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, and only as the
 * checker.  The product has no CPU path.
 *
 * Parity status: PINNED.  The reference ships no golden vectors of its own
 * (SURVEY.md F8), so the pin is the reference program itself:
 * oracle/Makefile compiles /root/reference/quantum_simulator.c unmodified into
 * oracle/_ref/ and tests/test_oracle.py demands bit-identical amplitudes
 * between this file and that build on the two shipped circuits and on seeded
 * random circuits over the reference gate set; tests/golden/ holds outputs of
 * that build so the pin also holds where /root/reference is absent.
 *
 * What is restated (reference = /root/reference/quantum_simulator.c):
 *   oc_apply_1q        <- execute_single_qubit_gate   :81-92
 *   oc_apply_cx        <- execute_cnot                :94-106
 *   oc_gate_matrix     <- the gate constant table     :184-211  (PI at :9)
 *   oc_run_file        <- compute_state_vector        :115-254  (grammar)
 *   oc_cdf             <- compute_state_cumulative_distribution :256-268
 *   oc_measure         <- measurement                 :270-283
 * Conventions kept: little-endian qubits (q[k] <-> bit k, :83), cx first
 * operand is the control (:229-235), rz(theta) is the PHASE gate
 * diag(1, e^{i theta}) (:205-208), all arithmetic in IEEE double.
 *
 * Extension (not in the reference, needed for BASELINE.json's synthetic
 * circuits): y, p, rx, ry, cz, cp, swap, ccx and arbitrary control masks.
 * These are validated against the reference build through exact identities
 * (RX = e^{-i t/2} H P(t) H, CP = P.CX.P.CX.P, SWAP = 3 CX) in
 * tests/test_oracle.py.
 */
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define OC_PI (2.0 * asin(1.0)) /* reference :9 */

/* State: interleaved (re, im) doubles -- layout-identical to C99
 * `double complex v[]` used by the reference (:125). */

/* out0 = a*m00 + b*m01 ; out1 = a*m10 + b*m11, m row-major (re,im) pairs.
 * The reference indexes its U[] as out0 = a*U[0] + b*U[2] (:88-89); every
 * matrix it builds is symmetric, so for its gate set U[2]==m01 and the two
 * conventions produce the same bits.  Products are expanded exactly as the
 * compiler expands C99 complex multiply for finite operands. */
void oc_apply_1q(double *v, int num_q, const double m[8], int target,
                 uint64_t controls)
{
    const uint64_t n = 1ULL << num_q, mask = 1ULL << target;
    for (uint64_t i = 0; i < n; i++) {
        if (i & mask) continue;               /* visit each pair once (:86) */
        if ((i & controls) != controls) continue;
        const uint64_t j = i | mask;
        const double ar = v[2 * i], ai = v[2 * i + 1];
        const double br = v[2 * j], bi = v[2 * j + 1];
        v[2 * i]     = (ar * m[0] - ai * m[1]) + (br * m[2] - bi * m[3]);
        v[2 * i + 1] = (ar * m[1] + ai * m[0]) + (br * m[3] + bi * m[2]);
        v[2 * j]     = (ar * m[4] - ai * m[5]) + (br * m[6] - bi * m[7]);
        v[2 * j + 1] = (ar * m[5] + ai * m[4]) + (br * m[7] + bi * m[6]);
    }
}

/* Conditional swap of the target pair where every control bit is set. */
void oc_apply_cx(double *v, int num_q, uint64_t controls, int target)
{
    const uint64_t n = 1ULL << num_q, mask = 1ULL << target;
    for (uint64_t i = 0; i < n; i++) {
        if ((i & mask) || (i & controls) != controls) continue;
        const uint64_t j = i | mask;
        double tr = v[2 * i], ti = v[2 * i + 1];
        v[2 * i] = v[2 * j]; v[2 * i + 1] = v[2 * j + 1];
        v[2 * j] = tr;       v[2 * j + 1] = ti;
    }
}

static void set_m(double m[8], double r00, double i00, double r01, double i01,
                  double r10, double i10, double r11, double i11)
{
    m[0] = r00; m[1] = i00; m[2] = r01; m[3] = i01;
    m[4] = r10; m[5] = i10; m[6] = r11; m[7] = i11;
}

/* cexp(I*x) for real x, as libm computes it: (cos x, sin x). */
static void phase(double x, double *re, double *im) { *re = cos(x); *im = sin(x); }

/* Gate table.  Returns 0 = single-qubit matrix in m, 1 = not a 1q name. */
int oc_gate_matrix(const char *name, double arg, double m[8])
{
    double pr, pi;
    if (!strcmp(name, "x")) { set_m(m, 0,0, 1,0, 1,0, 0,0); return 0; }
    if (!strcmp(name, "y")) { set_m(m, 0,0, 0,-1, 0,1, 0,0); return 0; }
    if (!strcmp(name, "z")) { set_m(m, 1,0, 0,0, 0,0, -1,0); return 0; }
    if (!strcmp(name, "h")) {
        double s = 1.0 / sqrt(2.0);                     /* :209-211 */
        set_m(m, s,0, s,0, s,0, -s,0); return 0;
    }
    if (!strcmp(name, "sx")) { set_m(m, .5,.5, .5,-.5, .5,-.5, .5,.5); return 0; } /* :190-192 */
    if (!strcmp(name, "s"))   { phase(OC_PI / 2.0, &pr, &pi);  set_m(m, 1,0, 0,0, 0,0, pr,pi); return 0; }
    if (!strcmp(name, "sdg")) { phase(-OC_PI / 2.0, &pr, &pi); set_m(m, 1,0, 0,0, 0,0, pr,pi); return 0; }
    if (!strcmp(name, "t"))   { phase(OC_PI / 4.0, &pr, &pi);  set_m(m, 1,0, 0,0, 0,0, pr,pi); return 0; }
    if (!strcmp(name, "tdg")) { phase(-OC_PI / 4.0, &pr, &pi); set_m(m, 1,0, 0,0, 0,0, pr,pi); return 0; }
    if (!strcmp(name, "rz") || !strcmp(name, "p")) {     /* phase gate, :205-208 */
        phase(arg, &pr, &pi); set_m(m, 1,0, 0,0, 0,0, pr,pi); return 0;
    }
    if (!strcmp(name, "rx")) {
        double c = cos(arg / 2.0), s = sin(arg / 2.0);
        set_m(m, c,0, 0,-s, 0,-s, c,0); return 0;
    }
    if (!strcmp(name, "ry")) {
        double c = cos(arg / 2.0), s = sin(arg / 2.0);
        set_m(m, c,0, -s,0, s,0, c,0); return 0;
    }
    return 1;
}

/* ---- grammar (reference :133-159, :162-181, :225-242) ------------------ */

static int is_sep(int c) /* what the reference skips between statements */
{
    return c == ' ' || c == '\t' || c == '\n' || c == ',' || c == ';' || !isgraph(c);
}

/* read next operand index: scan forward to '[' or '$', then an integer */
static int next_index(FILE *f, int *out)
{
    int c;
    while ((c = fgetc(f)) != EOF && c != '[' && c != '$') {}
    if (c == EOF) return -1;
    return fscanf(f, "%d", out) == 1 ? 0 : -1;
}

/*
 * Parse-and-execute, like the reference: statements are applied as they are
 * read.  *state is malloc'ed (caller frees).  Returns 0, or <0 on error:
 * -1 cannot open, -2 unknown token, -3 malformed, -4 out of memory.
 */
int oc_run_file(const char *path, double **state, int *num_q_out)
{
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int c, nq = 0, rc = 0;
    double *v = NULL;

    /* the two header statements (OPENQASM ...; include ...;) :133-141 */
    for (int k = 0; k < 2; k++)
        while ((c = fgetc(f)) != EOF && c != ';') {}

    for (;;) {
        while ((c = fgetc(f)) != EOF && (is_sep(c) || c == ']')) {}
        if (c == EOF) break;
        char tok[64]; int len = 0;
        while (c != EOF && isgraph(c) && c != '[' && len < 63) { tok[len++] = (char)c; c = fgetc(f); }
        tok[len] = 0;
        if (c == '[') ungetc(c, f);

        if (!strcmp(tok, "qubit")) {                        /* :162-181 */
            if (next_index(f, &nq) || nq < 0 || nq > 40) { rc = -3; break; }
            free(v);
            v = (double *)calloc((size_t)2 << nq, sizeof(double));
            if (!v) { rc = -4; break; }
            v[0] = 1.0;
            while ((c = fgetc(f)) != EOF && c != '\n') {}
            continue;
        }
        if (!v) { rc = -3; break; }

        /* split "name(arg)" */
        char name[64]; double arg = 0.0;
        strcpy(name, tok);
        char *par = strchr(name, '(');
        if (par) { *par = 0; if (sscanf(par + 1, "%lf", &arg) != 1) { rc = -3; break; } }

        int q0, q1, q2;
        double m[8];
        if (!strcmp(name, "cx") || !strcmp(name, "cz") || !strcmp(name, "cp") ||
            !strcmp(name, "swap")) {
            if (next_index(f, &q0) || next_index(f, &q1)) { rc = -3; break; }
            if (!strcmp(name, "cx")) oc_apply_cx(v, nq, 1ULL << q0, q1);
            else if (!strcmp(name, "swap")) {
                oc_apply_cx(v, nq, 1ULL << q0, q1);
                oc_apply_cx(v, nq, 1ULL << q1, q0);
                oc_apply_cx(v, nq, 1ULL << q0, q1);
            } else {
                oc_gate_matrix(!strcmp(name, "cz") ? "z" : "p", arg, m);
                oc_apply_1q(v, nq, m, q1, 1ULL << q0);
            }
        } else if (!strcmp(name, "ccx")) {
            if (next_index(f, &q0) || next_index(f, &q1) || next_index(f, &q2)) { rc = -3; break; }
            oc_apply_cx(v, nq, (1ULL << q0) | (1ULL << q1), q2);
        } else if (!oc_gate_matrix(name, arg, m)) {
            if (next_index(f, &q0)) { rc = -3; break; }
            if (!strcmp(name, "x")) oc_apply_cx(v, nq, 0, q0); /* X == swap; same bits as the matrix form */
            else oc_apply_1q(v, nq, m, q0, 0);
        } else { rc = -2; break; }
    }
    fclose(f);
    if (rc) { free(v); return rc; }
    *state = v; *num_q_out = nq;
    return 0;
}

/* inclusive prefix sum of |v|^2, serial, as the reference (:256-268);
 * cabs(z)*cabs(z) there == hypot()^2 here. */
void oc_cdf(const double *v, int num_q, double *out)
{
    double acc = 0.0;
    for (uint64_t i = 0; i < (1ULL << num_q); i++) {
        double a = hypot(v[2 * i], v[2 * i + 1]);
        acc += a * a;
        out[i] = acc;
    }
}

/* first index whose CDF entry is non-zero and >= r (:277-281) */
uint64_t oc_measure(const double *cdf, int num_q, double r)
{
    uint64_t idx = 0, last = (1ULL << num_q) - 1;
    while ((cdf[idx] == 0.0 || cdf[idx] < r) && idx < last) idx++;
    return idx;
}

void oc_free(void *p) { free(p); }
