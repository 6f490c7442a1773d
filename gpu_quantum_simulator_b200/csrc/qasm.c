/*
 * qasm.c -- circuit front end of libqsim_b200 (host, plain C).
 *
 * Produces the gate IR (qsb_gate_t[]) the device path consumes.  It accepts
 * everything the reference's streaming tokenizer accepts
 * (/root/reference/quantum_simulator.c:133-242) and the bare
 * "<num_q> <num_g>" header of the CUDA variants
 * (/root/reference/quantum_simulator_naive.cu:239-240), plus a documented
 * superset.  Unlike the reference it parses the whole file first and never
 * touches the state: execution is the device's job.
 *
 * Reference behaviours kept on purpose:
 *   - the register name of an operand is ignored; only the index after '['
 *     or '$' counts (:225-227)
 *   - `qubit[n] q;` and `qubit q[n];` both declare n qubits (:162-166)
 *   - rz(theta) == diag(1, e^{i theta}) (:205-208); constants use
 *     PI = 2*asin(1) and (cos, sin) of the angle exactly as cexp() does
 *   - cx: first operand is the control (:229-235)
 *   - an unknown gate name is an error: "Unknown token: <name>" (:213)
 */
#include <ctype.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qsim_b200.h"
#include "qsb_internal.h"

#define Q_PI (2.0 * asin(1.0))

/* ------------------------------------------------------------------ errors */
static __thread char g_err[512] = "";
void qsb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
const char *qsb_last_error(void) { return g_err; }
void qsb_free(void *p) { free(p); }

/* --------------------------------------------------------------- gate table */
static void mat(double m[8], double a, double b, double c, double d, double e,
                double f, double g, double h)
{
    m[0] = a; m[1] = b; m[2] = c; m[3] = d; m[4] = e; m[5] = f; m[6] = g; m[7] = h;
}

static void phase_mat(double m[8], double ang) { mat(m, 1, 0, 0, 0, 0, 0, cos(ang), sin(ang)); }

/* 0 on success; fills m and *np (params consumed). */
static int matrix_1q(const char *g, const double *p, int np, double m[8])
{
    const double r = 1.0 / sqrt(2.0);
#define NEED(k) do { if (np < (k)) { qsb_set_error("gate %s needs %d parameter(s)", g, (k)); return -1; } } while (0)
    if (!strcmp(g, "x")) mat(m, 0, 0, 1, 0, 1, 0, 0, 0);
    else if (!strcmp(g, "y")) mat(m, 0, 0, 0, -1, 0, 1, 0, 0);
    else if (!strcmp(g, "z")) mat(m, 1, 0, 0, 0, 0, 0, -1, 0);
    else if (!strcmp(g, "h")) mat(m, r, 0, r, 0, r, 0, -r, 0);
    else if (!strcmp(g, "id")) mat(m, 1, 0, 0, 0, 0, 0, 1, 0);
    else if (!strcmp(g, "sx")) mat(m, .5, .5, .5, -.5, .5, -.5, .5, .5);
    else if (!strcmp(g, "sxdg")) mat(m, .5, -.5, .5, .5, .5, .5, .5, -.5);
    else if (!strcmp(g, "s")) phase_mat(m, Q_PI / 2.0);
    else if (!strcmp(g, "sdg")) phase_mat(m, -Q_PI / 2.0);
    else if (!strcmp(g, "t")) phase_mat(m, Q_PI / 4.0);
    else if (!strcmp(g, "tdg")) phase_mat(m, -Q_PI / 4.0);
    else if (!strcmp(g, "rz") || !strcmp(g, "p") || !strcmp(g, "phase") || !strcmp(g, "u1")) {
        NEED(1); phase_mat(m, p[0]);
    } else if (!strcmp(g, "rx")) {
        NEED(1); double c = cos(p[0] / 2), s = sin(p[0] / 2);
        mat(m, c, 0, 0, -s, 0, -s, c, 0);
    } else if (!strcmp(g, "ry")) {
        NEED(1); double c = cos(p[0] / 2), s = sin(p[0] / 2);
        mat(m, c, 0, -s, 0, s, 0, c, 0);
    } else if (!strcmp(g, "u") || !strcmp(g, "u3") || !strcmp(g, "U")) {
        NEED(3); double c = cos(p[0] / 2), s = sin(p[0] / 2);
        mat(m, c, 0, -cos(p[2]) * s, -sin(p[2]) * s, cos(p[1]) * s, sin(p[1]) * s,
            cos(p[1] + p[2]) * c, sin(p[1] + p[2]) * c);
    } else return 1; /* not a 1q name */
#undef NEED
    return 0;
}

static int check_q(int q) { if (q < 0 || q > 62) { qsb_set_error("qubit index %d out of range", q); return -1; } return 0; }

int qsb_gate_from_name(const char *name, const double *params, int nparams,
                       const int *qubits, int nqubits, qsb_gate_t *out, int *nout)
{
    if (!name || !out || !nout || (nqubits > 0 && !qubits)) { qsb_set_error("qsb_gate_from_name: null argument"); return QSB_ERR_ARG; }
    for (int i = 0; i < nqubits; i++) {
        if (check_q(qubits[i])) return QSB_ERR_ARG;
        for (int j = 0; j < i; j++)
            if (qubits[i] == qubits[j]) { qsb_set_error("gate %s: repeated operand %d", name, qubits[i]); return QSB_ERR_PARSE; }
    }
    char base[32];
    int nctrl = 0;
    size_t len = strlen(name);
    if (len == 0 || len >= sizeof base) { qsb_set_error("Unknown token: %s", name); return QSB_ERR_PARSE; }

    if (!strcmp(name, "swap")) { /* three CX, as the reference's gate set would spell it */
        if (nqubits != 2) { qsb_set_error("swap needs 2 operands"); return QSB_ERR_PARSE; }
        for (int k = 0; k < 3; k++) {
            memset(&out[k], 0, sizeof out[k]);
            int c = qubits[k & 1], t = qubits[(k & 1) ^ 1];
            out[k].controls = 1ULL << c; out[k].target = t;
            mat(out[k].m, 0, 0, 1, 0, 1, 0, 0, 0);
        }
        *nout = 3;
        return QSB_OK;
    }
    /* strip control prefixes: cx, ccx, cz, cp, ch, cy, crx ... */
    const char *b = name;
    double m[8];
    int rc = matrix_1q(b, params, nparams, m);
    while (rc == 1 && *b == 'c' && b[1]) { b++; nctrl++; rc = matrix_1q(b, params, nparams, m); }
    if (rc == 1 && !strcmp(name, "cnot")) { nctrl = 1; rc = matrix_1q("x", params, nparams, m); }
    if (rc == 1 && (!strcmp(name, "toffoli"))) { nctrl = 2; rc = matrix_1q("x", params, nparams, m); }
    if (rc == 1) { qsb_set_error("Unknown token: %s", name); return QSB_ERR_PARSE; }
    if (rc < 0) return QSB_ERR_PARSE;
    (void)base;
    if (nqubits != nctrl + 1) { qsb_set_error("gate %s needs %d operand(s), got %d", name, nctrl + 1, nqubits); return QSB_ERR_PARSE; }
    memset(out, 0, sizeof *out);
    for (int i = 0; i < nctrl; i++) out->controls |= 1ULL << qubits[i];
    out->target = qubits[nctrl];
    memcpy(out->m, m, sizeof m);
    *nout = 1;
    return QSB_OK;
}

/* ------------------------------------------------------ expression evaluator
 * number | pi | (expr) | -x | x+y | x-y | x*y | x/y     (for gate parameters) */
typedef struct { const char *s; int err; } ex_t;
static void ex_ws(ex_t *e) { while (*e->s && isspace((unsigned char)*e->s)) e->s++; }
static double ex_sum(ex_t *e);
static double ex_atom(ex_t *e)
{
    ex_ws(e);
    if (*e->s == '(') { e->s++; double v = ex_sum(e); ex_ws(e); if (*e->s == ')') e->s++; else e->err = 1; return v; }
    if (*e->s == '-') { e->s++; return -ex_atom(e); }
    if (*e->s == '+') { e->s++; return ex_atom(e); }
    if (!strncmp(e->s, "pi", 2) && !isalnum((unsigned char)e->s[2])) { e->s += 2; return Q_PI; }
    if (!strncmp(e->s, "\xcf\x80", 2)) { e->s += 2; return Q_PI; }
    char *end; double v = strtod(e->s, &end);
    if (end == e->s) { e->err = 1; return 0; }
    e->s = end; return v;
}
static double ex_prod(ex_t *e)
{
    double v = ex_atom(e);
    for (;;) {
        ex_ws(e);
        if (*e->s == '*') { e->s++; v *= ex_atom(e); }
        else if (*e->s == '/') { e->s++; v /= ex_atom(e); }
        else return v;
    }
}
static double ex_sum(ex_t *e)
{
    double v = ex_prod(e);
    for (;;) {
        ex_ws(e);
        if (*e->s == '+') { e->s++; v += ex_prod(e); }
        else if (*e->s == '-') { e->s++; v -= ex_prod(e); }
        else return v;
    }
}

/* ------------------------------------------------------------------- parser */
typedef struct { qsb_gate_t *g; size_t n, cap; } gvec_t;
static int gv_push(gvec_t *v, const qsb_gate_t *g, int k)
{
    if (v->n + (size_t)k > v->cap) {
        size_t nc = v->cap ? v->cap * 2 : 1024;
        while (nc < v->n + (size_t)k) nc *= 2;
        qsb_gate_t *ng = (qsb_gate_t *)realloc(v->g, nc * sizeof *ng);
        if (!ng) { qsb_set_error("Malloc error"); return -1; }
        v->g = ng; v->cap = nc;
    }
    memcpy(v->g + v->n, g, (size_t)k * sizeof *g);
    v->n += (size_t)k;
    return 0;
}

int qsb_parse_qasm_string(const char *text, int *num_qubits, qsb_gate_t **gates, size_t *n)
{
    if (!text || !num_qubits || !gates || !n) { qsb_set_error("qsb_parse_qasm_string: null argument"); return QSB_ERR_ARG; }
    const char *s = text;
    gvec_t gv = {0, 0, 0};
    int nq = -1, max_q = -1;

    /* CUDA-variant header: "<num_q> <num_g>" */
    while (*s && isspace((unsigned char)*s)) s++;
    if (isdigit((unsigned char)*s)) {
        char *e1, *e2;
        long a = strtol(s, &e1, 10);
        long b = strtol(e1, &e2, 10);
        if (e2 == e1) { qsb_set_error("bad \"<num_q> <num_g>\" header"); return QSB_ERR_PARSE; }
        (void)b; /* the gate count is implied by the text */
        nq = (int)a; s = e2;
    }

    while (*s) {
        /* separators the reference skips (:147-149, :240-242) */
        while (*s && (isspace((unsigned char)*s) || *s == ',' || *s == ';' || *s == ']' || !isgraph((unsigned char)*s))) s++;
        if (!*s) break;
        if (s[0] == '/' && s[1] == '/') { while (*s && *s != '\n') s++; continue; }
        if (s[0] == '/' && s[1] == '*') { const char *e = strstr(s + 2, "*/"); s = e ? e + 2 : s + strlen(s); continue; }

        /* name: graph chars up to '[' '(' ';' or blank (:150-159) */
        char name[64]; int len = 0;
        while (*s && isgraph((unsigned char)*s) && *s != '[' && *s != '(' && *s != ';' && *s != '$' && len < 63) name[len++] = *s++;
        name[len] = 0;
        if (len == 0) { s++; continue; }

        { /* classical assignment (`c = measure q;`, `c[0] = measure q[0];`): ignored */
            const char *e = s; int is_assign = 0;
            while (*e && *e != ';' && *e != '\n') { if (*e == '=') { is_assign = 1; break; } e++; }
            if (is_assign) { while (*s && *s != ';' && *s != '\n') s++; continue; }
        }
        if (!strcmp(name, "OPENQASM") || !strcmp(name, "include") || !strcmp(name, "barrier") ||
            !strcmp(name, "measure") || !strcmp(name, "bit") || !strcmp(name, "creg") ||
            !strcmp(name, "reset") || !strcmp(name, "gphase")) {
            while (*s && *s != ';' && *s != '\n') s++; /* whole statement ignored */
            continue;
        }
        if (!strcmp(name, "qubit") || !strcmp(name, "qreg")) {
            while (*s && *s != '[' && *s != '$' && *s != ';' && *s != '\n') s++;
            if (*s != '[' && *s != '$') { nq = 1; continue; } /* `qubit q;` */
            s++;
            char *e; long v = strtol(s, &e, 10);
            if (e == s || v < 1 || v > 62) { qsb_set_error("bad qubit count in declaration"); free(gv.g); return QSB_ERR_PARSE; }
            if (nq >= 0 && gv.n) { qsb_set_error("only a single quantum register is supported"); free(gv.g); return QSB_ERR_PARSE; }
            nq = (int)v; s = e;
            while (*s && *s != ';' && *s != '\n') s++;
            continue;
        }

        /* parameters */
        double par[4]; int np = 0;
        while (*s == ' ' || *s == '\t') s++;
        if (*s == '(') {
            int depth = 0; const char *st = s + 1; const char *p = s;
            for (; *p; p++) {
                if (*p == '(') depth++;
                else if (*p == ')') { if (--depth == 0) break; }
            }
            if (!*p) { qsb_set_error("gate %s: unterminated '('", name); free(gv.g); return QSB_ERR_PARSE; }
            char buf[256]; size_t bl = (size_t)(p - st);
            if (bl >= sizeof buf) { qsb_set_error("gate %s: parameter list too long", name); free(gv.g); return QSB_ERR_PARSE; }
            memcpy(buf, st, bl); buf[bl] = 0;
            char *tokp = buf;
            while (tokp && *tokp && np < 4) {
                /* split on top-level commas */
                int d = 0; char *q = tokp;
                for (; *q; q++) { if (*q == '(') d++; else if (*q == ')') d--; else if (*q == ',' && d == 0) break; }
                char save = *q; *q = 0;
                ex_t ex = { tokp, 0 };
                double v = ex_sum(&ex); ex_ws(&ex);
                if (ex.err || *ex.s) { qsb_set_error("gate %s: cannot parse parameter \"%s\"", name, tokp); free(gv.g); return QSB_ERR_PARSE; }
                par[np++] = v;
                tokp = save ? q + 1 : NULL;
            }
            s = p + 1;
        }

        /* operands up to end of statement: every '[' or '$' introduces an index (:225-233) */
        int ops[8], nops = 0;
        while (*s && *s != ';' && *s != '\n') {
            if (*s == '[' || *s == '$') {
                char *e; long v = strtol(s + 1, &e, 10);
                if (e == s + 1) { qsb_set_error("gate %s: bad operand", name); free(gv.g); return QSB_ERR_PARSE; }
                if (nops < 8) ops[nops++] = (int)v;
                s = e;
            } else s++;
        }
        qsb_gate_t tmp[3]; int k = 0;
        int rc = qsb_gate_from_name(name, par, np, ops, nops, tmp, &k);
        if (rc) { free(gv.g); return rc; }
        for (int i = 0; i < nops; i++) if (ops[i] > max_q) max_q = ops[i];
        if (gv_push(&gv, tmp, k)) { free(gv.g); return QSB_ERR_NOMEM; }
    }
    if (nq < 0) { qsb_set_error("no qubit declaration found"); free(gv.g); return QSB_ERR_PARSE; }
    if (max_q >= nq) { qsb_set_error("operand q[%d] exceeds the declared %d qubits", max_q, nq); free(gv.g); return QSB_ERR_PARSE; }
    *num_qubits = nq; *gates = gv.g; *n = gv.n;
    if (!gv.g) *gates = (qsb_gate_t *)calloc(1, sizeof(qsb_gate_t));
    return QSB_OK;
}

int qsb_parse_qasm_file(const char *path, int *num_qubits, qsb_gate_t **gates, size_t *n)
{
    if (!path) { qsb_set_error("qsb_parse_qasm_file: null path"); return QSB_ERR_ARG; }
    FILE *f = fopen(path, "rb");
    if (!f) { qsb_set_error("ERROR: cannot open circuit file"); return QSB_ERR_IO; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *buf = (char *)malloc((size_t)sz + 1);
    if (!buf) { fclose(f); qsb_set_error("Malloc error"); return QSB_ERR_NOMEM; }
    size_t got = fread(buf, 1, (size_t)sz, f);
    fclose(f);
    buf[got] = 0;
    int rc = qsb_parse_qasm_string(buf, num_qubits, gates, n);
    free(buf);
    return rc;
}
