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
 * Superset: more gates and c-prefixes, `gate` definitions (expanded on use),
 * ctrl @ / ctrl(n) @ / negctrl @ / inv @ / pow(k) @, gphase, several registers (names then matter),
 * whole-register operands, parameter expressions with pi / tau / euler and sin cos tan exp ln sqrt.
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
 * number | pi | tau | euler | name (a parameter of the enclosing `gate`) | f(expr) for f in
 * sin cos tan exp ln sqrt | (expr) | -x | x+y | x-y | x*y | x/y | x**y        (gate parameters) */
typedef struct { char name[8][24]; double val[8]; int n; } pbind_t;
typedef struct { const char *s; int err; const pbind_t *b; } ex_t;
static void ex_ws(ex_t *e) { while (*e->s && isspace((unsigned char)*e->s)) e->s++; }
static double ex_sum(ex_t *e);
static double ex_atom(ex_t *e)
{
    ex_ws(e);
    if (*e->s == '(') { e->s++; double v = ex_sum(e); ex_ws(e); if (*e->s == ')') e->s++; else e->err = 1; return v; }
    if (*e->s == '-') { e->s++; return -ex_atom(e); }
    if (*e->s == '+') { e->s++; return ex_atom(e); }
    if (!strncmp(e->s, "\xcf\x80", 2)) { e->s += 2; return Q_PI; }
    if (isalpha((unsigned char)*e->s) || *e->s == '_') {
        char id[24]; int n = 0;
        while ((isalnum((unsigned char)*e->s) || *e->s == '_') && n < 23) id[n++] = *e->s++;
        id[n] = 0;
        if (!strcmp(id, "pi")) return Q_PI;
        if (!strcmp(id, "tau")) return 2.0 * Q_PI;
        if (!strcmp(id, "euler")) return exp(1.0);
        ex_ws(e);
        if (*e->s == '(') {
            double (*f)(double) = !strcmp(id, "sin") ? sin : !strcmp(id, "cos") ? cos : !strcmp(id, "tan") ? tan :
                                  !strcmp(id, "exp") ? exp : !strcmp(id, "ln") ? log : !strcmp(id, "sqrt") ? sqrt : NULL;
            if (!f) { e->err = 1; return 0; }
            e->s++; double v = ex_sum(e); ex_ws(e);
            if (*e->s == ')') e->s++; else e->err = 1;
            return f(v);
        }
        if (e->b) for (int i = 0; i < e->b->n; i++) if (!strcmp(id, e->b->name[i])) return e->b->val[i];
        e->err = 1; return 0;
    }
    char *end; double v = strtod(e->s, &end);
    if (end == e->s) { e->err = 1; return 0; }
    e->s = end; return v;
}
static double ex_pow(ex_t *e)
{
    double v = ex_atom(e);
    ex_ws(e);
    if (e->s[0] == '*' && e->s[1] == '*') { e->s += 2; return pow(v, ex_pow(e)); }
    return v;
}
static double ex_prod(ex_t *e)
{
    double v = ex_pow(e);
    for (;;) {
        ex_ws(e);
        if (*e->s == '*' && e->s[1] != '*') { e->s++; v *= ex_pow(e); }
        else if (*e->s == '/') { e->s++; v /= ex_pow(e); }
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

/* ------------------------------------------------------------------- parser
 * Superset of the reference grammar (see the header comment).  Beyond the reference:
 *   gate NAME(p, ...) a, b { body }      user gates, expanded on use (nested, parameters are expressions)
 *   ctrl @ / ctrl(n) @ / negctrl @ / inv @ / pow(k) @     gate modifiers (k a non-negative integer)
 *   gphase(theta);                        global phase (a phase gate on its controls under ctrl @)
 *   several quantum registers, whole-register operands (`h q;` applies the gate to every qubit)      */
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

typedef struct { char name[32]; int off, size; } qreg_t;
typedef struct { char name[32]; int npar, nq; char par[8][24], qarg[8][24]; const char *body, *body_end; } gdef_t;
typedef struct { char name[8][24]; int idx[8]; int n; } qbind_t;
typedef struct {
    gvec_t gv;
    qreg_t regs[16]; int nregs;
    int nq, max_q;
    gdef_t *defs; int ndefs;
} pctx_t;
#define QSB_MAX_DEFS 128
#define PFAIL(...) do { qsb_set_error(__VA_ARGS__); return QSB_ERR_PARSE; } while (0)

static int is_id_start(int c) { return isalpha(c) || c == '_'; }
static const char *skip_ws(const char *s, const char *end) { while (s < end && (isspace((unsigned char)*s))) s++; return s; }
static const char *read_id(const char *s, const char *end, char *out, int cap)
{
    int n = 0;
    while (s < end && (isalnum((unsigned char)*s) || *s == '_') && n < cap - 1) out[n++] = *s++;
    out[n] = 0;
    return s;
}

/* adjoint of a (controlled) one-qubit gate */
static void gate_adjoint(qsb_gate_t *g)
{
    double t[8];
    t[0] = g->m[0]; t[1] = -g->m[1]; t[2] = g->m[4]; t[3] = -g->m[5];
    t[4] = g->m[2]; t[5] = -g->m[3]; t[6] = g->m[6]; t[7] = -g->m[7];
    memcpy(g->m, t, sizeof t);
}

static int parse_block(pctx_t *c, const char *s, const char *end, const pbind_t *pb, const qbind_t *qb, int depth);

/* apply `name(par) ops` with nctrl extra leading control operands, negated controls in negmask (bit i = i-th
 * control operand), inverse and integer power; appends to c->gv */
static int apply_gate(pctx_t *c, const char *name, const double *par, int np, const int *ops, int nops,
                      int nctrl, unsigned negmask, int inv, int power, int depth)
{
    if (depth > 32) PFAIL("gate definitions nested too deeply (recursive gate?)");
    if (nctrl > nops) PFAIL("gate %s: %d control operand(s) requested, %d operand(s) given", name, nctrl, nops);
    uint64_t cmask = 0;
    for (int i = 0; i < nctrl; i++) {
        if (check_q(ops[i])) return QSB_ERR_ARG;
        if ((cmask >> ops[i]) & 1) PFAIL("gate %s: repeated operand %d", name, ops[i]);
        cmask |= 1ULL << ops[i];
    }
    const int *tops = ops + nctrl; const int ntops = nops - nctrl;
    gvec_t tmp = {0, 0, 0};
    int rc = QSB_OK;
    if (!strcmp(name, "gphase")) {
        if (np != 1) { free(tmp.g); PFAIL("gphase needs 1 parameter"); }
        if (ntops != 0) { free(tmp.g); PFAIL("gphase takes no operands"); }
        qsb_gate_t g; memset(&g, 0, sizeof g);
        if (nctrl == 0) { g.target = 0; mat(g.m, cos(par[0]), sin(par[0]), 0, 0, 0, 0, cos(par[0]), sin(par[0])); }
        else {   /* ctrl @ gphase(t) c...  ==  phase gate on the last control, controlled by the others */
            g.target = ops[nctrl - 1]; g.controls = cmask & ~(1ULL << g.target);
            phase_mat(g.m, par[0]);
            cmask = 0;
        }
        if (gv_push(&tmp, &g, 1)) return QSB_ERR_NOMEM;
    } else {
        const gdef_t *d = NULL;
        for (int i = c->ndefs - 1; i >= 0 && !d; i--) if (!strcmp(c->defs[i].name, name)) d = &c->defs[i];
        if (d) {
            if (np != d->npar) { PFAIL("gate %s needs %d parameter(s), got %d", name, d->npar, np); }
            if (ntops != d->nq) { PFAIL("gate %s needs %d operand(s), got %d", name, d->nq, ntops); }
            pbind_t pbn; qbind_t qbn; memset(&pbn, 0, sizeof pbn); memset(&qbn, 0, sizeof qbn);
            for (int i = 0; i < np; i++) { strcpy(pbn.name[i], d->par[i]); pbn.val[i] = par[i]; }
            pbn.n = np;
            for (int i = 0; i < ntops; i++) {
                for (int j = 0; j < i; j++) if (tops[j] == tops[i]) PFAIL("gate %s: repeated operand %d", name, tops[i]);
                strcpy(qbn.name[i], d->qarg[i]); qbn.idx[i] = tops[i];
            }
            qbn.n = ntops;
            gvec_t save = c->gv; c->gv = tmp;
            rc = parse_block(c, d->body, d->body_end, &pbn, &qbn, depth + 1);
            tmp = c->gv; c->gv = save;
            if (rc) { free(tmp.g); return rc; }
        } else {
            qsb_gate_t g3[3]; int k = 0;
            rc = qsb_gate_from_name(name, par, np, tops, ntops, g3, &k);
            if (rc) return rc;
            if (gv_push(&tmp, g3, k)) return QSB_ERR_NOMEM;
        }
    }
    /* modifiers: inverse = reversed order of adjoints; power = repetition; controls = every gate controlled */
    if (inv) {
        for (size_t i = 0; i < tmp.n / 2; i++) { qsb_gate_t t = tmp.g[i]; tmp.g[i] = tmp.g[tmp.n - 1 - i]; tmp.g[tmp.n - 1 - i] = t; }
        for (size_t i = 0; i < tmp.n; i++) gate_adjoint(&tmp.g[i]);
    }
    for (size_t i = 0; i < tmp.n; i++) {
        qsb_gate_t *gi = &tmp.g[i];
        /* a global phase inside a gate body (gphase: scalar * identity on a dummy target) has no operand of its own:
         * under controls it becomes a phase gate on one control, controlled by the others */
        const int scalar = gi->controls == 0 && gi->m[2] == 0 && gi->m[3] == 0 && gi->m[4] == 0 && gi->m[5] == 0 &&
                           gi->m[0] == gi->m[6] && gi->m[1] == gi->m[7];
        if (scalar && cmask) {
            const double pr = gi->m[0], pi = gi->m[1];
            gi->target = __builtin_ctzll(cmask); gi->controls = cmask & ~(1ULL << gi->target);
            mat(gi->m, 1, 0, 0, 0, 0, 0, pr, pi);
            continue;
        }
        if ((gi->controls & cmask) || ((cmask >> gi->target) & 1)) { free(tmp.g); PFAIL("gate %s: a control is also an operand of the gate", name); }
        gi->controls |= cmask;
    }
    qsb_gate_t xg; memset(&xg, 0, sizeof xg); mat(xg.m, 0, 0, 1, 0, 1, 0, 0, 0);
    for (int i = 0; i < nctrl; i++) if ((negmask >> i) & 1) { xg.target = ops[i]; if (gv_push(&c->gv, &xg, 1)) { free(tmp.g); return QSB_ERR_NOMEM; } }
    for (int r = 0; r < power; r++) if (tmp.n && gv_push(&c->gv, tmp.g, (int)tmp.n)) { free(tmp.g); return QSB_ERR_NOMEM; }
    for (int i = 0; i < nctrl; i++) if ((negmask >> i) & 1) { xg.target = ops[i]; if (gv_push(&c->gv, &xg, 1)) { free(tmp.g); return QSB_ERR_NOMEM; } }
    free(tmp.g);
    for (int i = 0; i < nops; i++) if (ops[i] > c->max_q) c->max_q = ops[i];
    return QSB_OK;
}

/* parse "( expr, expr, ... )" starting at '(' ; returns the position after ')' */
static int parse_params(const char *gname, const char **ps, const char *end, const pbind_t *pb, double *par, int *np)
{
    const char *s = *ps;
    int depth = 0; const char *st = s + 1; const char *p = s;
    for (; p < end; p++) { if (*p == '(') depth++; else if (*p == ')') { if (--depth == 0) break; } }
    if (p >= end) PFAIL("gate %s: unterminated '('", gname);
    char buf[512]; size_t bl = (size_t)(p - st);
    if (bl >= sizeof buf) PFAIL("gate %s: parameter list too long", gname);
    memcpy(buf, st, bl); buf[bl] = 0;
    char *tokp = buf; *np = 0;
    while (tokp && *tokp) {
        int d = 0; char *q = tokp;
        for (; *q; q++) { if (*q == '(') d++; else if (*q == ')') d--; else if (*q == ',' && d == 0) break; }
        char save = *q; *q = 0;
        ex_t ex = { tokp, 0, pb };
        double v = ex_sum(&ex); ex_ws(&ex);
        if (ex.err || *ex.s) PFAIL("gate %s: cannot parse parameter \"%s\"", gname, tokp);
        if (*np >= 8) PFAIL("gate %s: too many parameters", gname);
        par[(*np)++] = v;
        tokp = save ? q + 1 : NULL;
    }
    *ps = p + 1;
    return QSB_OK;
}

static int parse_block(pctx_t *c, const char *s, const char *end, const pbind_t *pb, const qbind_t *qb, int depth)
{
    while (s < end) {
        /* separators the reference skips (:147-149, :240-242) */
        while (s < end && (isspace((unsigned char)*s) || *s == ',' || *s == ';' || *s == ']' || !isgraph((unsigned char)*s))) s++;
        if (s >= end) break;
        if (s[0] == '/' && s + 1 < end && s[1] == '/') { while (s < end && *s != '\n') s++; continue; }
        if (s[0] == '/' && s + 1 < end && s[1] == '*') { const char *e = strstr(s + 2, "*/"); s = (e && e < end) ? e + 2 : end; continue; }

        /* modifiers and gate name */
        int nctrl = 0, inv = 0; long long power = 1; unsigned negmask = 0;
        char name[64];
        for (;;) {
            int len = 0;
            while (s < end && isgraph((unsigned char)*s) && *s != '[' && *s != '(' && *s != ';' && *s != '$' && *s != '@' && *s != '{' && len < 63) name[len++] = *s++;
            name[len] = 0;
            if (len == 0) break;
            const int is_mod = !strcmp(name, "ctrl") || !strcmp(name, "negctrl") || !strcmp(name, "inv") || !strcmp(name, "pow");
            if (!is_mod) break;
            /* a modifier must be followed by [ (n) ] @ ; otherwise it is an ordinary (unknown) token */
            const char *t = skip_ws(s, end);
            double marg = 1; int has_arg = 0;
            if (t < end && *t == '(') { double pv[8]; int npv = 0; int rc = parse_params(name, &t, end, pb, pv, &npv); if (rc) return rc; if (npv != 1) PFAIL("modifier %s takes one argument", name); marg = pv[0]; has_arg = 1; t = skip_ws(t, end); }
            if (t >= end || *t != '@') break;
            s = skip_ws(t + 1, end);
            if (!strcmp(name, "inv")) inv ^= 1;
            else if (!strcmp(name, "pow")) {
                if (!has_arg || marg < 0 || marg != floor(marg) || marg > 1e6) PFAIL("pow(k) @ needs a non-negative integer k");
                power *= (long long)marg;      /* 64-bit: stacked pow modifiers must not wrap */
                if (power > 1000000) PFAIL("pow modifiers multiply to more than 1e6 repetitions");
            } else {
                const int k = has_arg ? (int)marg : 1;
                if (k < 1 || nctrl + k > 8 || marg != floor(marg)) PFAIL("%s(n) @ needs a small positive integer n", name);
                if (!strcmp(name, "negctrl")) for (int i = 0; i < k; i++) negmask |= 1u << (nctrl + i);
                nctrl += k;
            }
        }
        if (name[0] == 0) { s++; continue; }

        { /* classical assignment (`c = measure q;`, `c[0] = measure q[0];`): ignored */
            const char *e = s; int is_assign = 0;
            while (e < end && *e != ';' && *e != '\n' && *e != '{') {
                if (e[0] == '/' && e + 1 < end && (e[1] == '/' || e[1] == '*')) break;   /* a comment is not part of the statement */
                if (*e == '=') { is_assign = 1; break; }
                e++;
            }
            if (is_assign) { while (s < end && *s != ';' && *s != '\n') s++; continue; }
        }
        if (!strcmp(name, "OPENQASM") || !strcmp(name, "include") || !strcmp(name, "barrier") ||
            !strcmp(name, "measure") || !strcmp(name, "bit") || !strcmp(name, "creg") || !strcmp(name, "reset")) {
            while (s < end && *s != ';' && *s != '\n') s++; /* whole statement ignored */
            continue;
        }
        if (!strcmp(name, "gate")) {
            if (depth > 0) PFAIL("gate definitions cannot be nested");
            if (c->ndefs >= QSB_MAX_DEFS) PFAIL("too many gate definitions");
            gdef_t *d = &c->defs[c->ndefs]; memset(d, 0, sizeof *d);
            s = skip_ws(s, end);
            s = read_id(s, end, d->name, (int)sizeof d->name);
            if (!d->name[0]) PFAIL("gate definition without a name");
            s = skip_ws(s, end);
            if (s < end && *s == '(') {
                s++;
                for (;;) {
                    s = skip_ws(s, end);
                    if (s < end && *s == ')') { s++; break; }
                    if (d->npar >= 8) PFAIL("gate %s: too many parameters", d->name);
                    const char *t = read_id(s, end, d->par[d->npar], 24);
                    if (t == s) PFAIL("gate %s: bad parameter list", d->name);
                    d->npar++; s = skip_ws(t, end);
                    if (s < end && *s == ',') s++;
                }
            }
            for (;;) {
                s = skip_ws(s, end);
                if (s >= end) PFAIL("gate %s: missing body", d->name);
                if (*s == '{') break;
                if (*s == ',') { s++; continue; }
                if (d->nq >= 8) PFAIL("gate %s: too many qubit arguments", d->name);
                const char *t = read_id(s, end, d->qarg[d->nq], 24);
                if (t == s) PFAIL("gate %s: bad qubit argument list", d->name);
                d->nq++; s = t;
            }
            const char *b = s + 1; int lvl = 1; const char *e = b;
            for (; e < end && lvl; e++) { if (*e == '{') lvl++; else if (*e == '}') lvl--; }
            if (lvl) PFAIL("gate %s: unterminated body", d->name);
            d->body = b; d->body_end = e - 1;
            c->ndefs++;
            s = e;
            continue;
        }
        if (!strcmp(name, "qubit") || !strcmp(name, "qreg")) {
            if (depth > 0) PFAIL("declaration inside a gate body");
            /* `qubit[n] name;`  `qubit name[n];`  `qubit name;`  (the reference scans to '[' or '$', :162-166) */
            char rname[32] = ""; long v = 1; int have_n = 0;
            while (s < end && *s != ';' && *s != '\n') {
                if (*s == '[' || *s == '$') { char *e; v = strtol(s + 1, &e, 10); if (e == s + 1 || v < 1 || v > 62) PFAIL("bad qubit count in declaration"); have_n = 1; s = e; }
                else if (is_id_start((unsigned char)*s) && !rname[0]) s = read_id(s, end, rname, (int)sizeof rname);
                else s++;
            }
            (void)have_n;
            if (c->nq < 0) c->nq = 0;
            if (c->nregs >= 16) PFAIL("too many quantum registers");
            if (c->nq + v > 62) PFAIL("bad qubit count in declaration");
            qreg_t *r = &c->regs[c->nregs++];
            strcpy(r->name, rname); r->off = c->nq; r->size = (int)v;
            c->nq += (int)v;
            continue;
        }

        /* parameters */
        double par[8]; int np = 0;
        while (s < end && (*s == ' ' || *s == '\t')) s++;
        if (s < end && *s == '(') { int rc = parse_params(name, &s, end, pb, par, &np); if (rc) return rc; }

        /* operands up to the end of the statement.  `$k` and `name[k]` are indices; a bare name is a qubit argument
         * of the enclosing gate or a whole register.  With a single register the register name is ignored, as in
         * the reference (:225-233). */
        int ops[16], nops = 0, whole[16], nwhole = 0, bsize = -1;
        int after_comma = 0;
        while (s < end && *s != ';') {
            if (*s == '\n') {   /* a newline ends the statement unless the operand list continues (`cx q[0],\n q[1];`) */
                if (after_comma) { s++; continue; }
                break;
            }
            if (s[0] == '/' && s + 1 < end && s[1] == '/') { while (s < end && *s != '\n') s++; continue; }
            if (s[0] == '/' && s + 1 < end && s[1] == '*') { const char *e = strstr(s + 2, "*/"); s = (e && e < end) ? e + 2 : end; continue; }
            if (*s == ',') { after_comma = 1; s++; continue; }
            if (!isspace((unsigned char)*s)) after_comma = 0;
            if (*s == '$') {
                char *e; long v = strtol(s + 1, &e, 10);
                if (e == s + 1) PFAIL("gate %s: bad operand", name);
                if (v < 0 || v > 62) PFAIL("gate %s: operand index %ld out of range", name, v);
                if (nops < 16) { whole[nops] = -1; ops[nops++] = (int)v; }
                s = e;
            } else if (*s == '[') {   /* index without a usable name in front: reference behaviour */
                char *e; long v = strtol(s + 1, &e, 10);
                if (e == s + 1) PFAIL("gate %s: bad operand", name);
                if (v < 0 || v > 62) PFAIL("gate %s: operand index %ld out of range", name, v);
                if (nops < 16) { whole[nops] = -1; ops[nops++] = (int)v; }
                s = e;
            } else if (is_id_start((unsigned char)*s)) {
                char id[32]; const char *t = read_id(s, end, id, (int)sizeof id);
                const char *u = skip_ws(t, end);
                int reg = -1;
                for (int i = 0; i < c->nregs; i++) if (!strcmp(c->regs[i].name, id)) reg = i;
                if (u < end && *u == '[') {
                    char *e; long v = strtol(u + 1, &e, 10);
                    if (e == u + 1) PFAIL("gate %s: bad operand", name);
                    if (reg >= 0 && c->nregs > 1) {
                        if (v < 0 || v >= c->regs[reg].size) PFAIL("operand %s[%ld] exceeds the register", id, v);
                        v += c->regs[reg].off;
                    }
                    if (v < 0 || v > 62) PFAIL("gate %s: operand index %ld out of range", name, v);
                    if (nops < 16) { whole[nops] = -1; ops[nops++] = (int)v; }
                    s = e;
                } else {
                    int bound = -1;
                    if (qb) for (int i = 0; i < qb->n; i++) if (!strcmp(qb->name[i], id)) bound = qb->idx[i];
                    if (bound >= 0) { if (nops < 16) { whole[nops] = -1; ops[nops++] = bound; } }
                    else if (reg >= 0 && depth == 0) {
                        if (bsize >= 0 && bsize != c->regs[reg].size) PFAIL("gate %s: whole-register operands of different sizes", name);
                        bsize = c->regs[reg].size;
                        if (nops < 16) { whole[nops] = reg; ops[nops++] = 0; nwhole++; }
                    } else PFAIL("gate %s: unknown operand %s", name, id);
                    s = t;
                }
            } else s++;
        }
        const int reps = nwhole ? bsize : 1;
        for (int r = 0; r < reps; r++) {
            int o2[16];
            for (int i = 0; i < nops; i++) o2[i] = whole[i] >= 0 ? c->regs[whole[i]].off + r : ops[i];
            int rc = apply_gate(c, name, par, np, o2, nops, nctrl, negmask, inv, (int)power, depth);
            if (rc) return rc;
        }
    }
    return QSB_OK;
}

int qsb_parse_qasm_string(const char *text, int *num_qubits, qsb_gate_t **gates, size_t *n)
{
    if (!text || !num_qubits || !gates || !n) { qsb_set_error("qsb_parse_qasm_string: null argument"); return QSB_ERR_ARG; }
    const char *s = text;
    pctx_t c; memset(&c, 0, sizeof c);
    c.nq = -1; c.max_q = -1;
    c.defs = (gdef_t *)calloc(QSB_MAX_DEFS, sizeof(gdef_t));
    if (!c.defs) { qsb_set_error("Malloc error"); return QSB_ERR_NOMEM; }

    /* CUDA-variant header: "<num_q> <num_g>" */
    while (*s && isspace((unsigned char)*s)) s++;
    if (isdigit((unsigned char)*s)) {
        char *e1, *e2;
        long a = strtol(s, &e1, 10);
        long b = strtol(e1, &e2, 10);
        if (e2 == e1) { free(c.defs); qsb_set_error("bad \"<num_q> <num_g>\" header"); return QSB_ERR_PARSE; }
        if (a < 1 || a > 62) { free(c.defs); qsb_set_error("bad qubit count %ld in the \"<num_q> <num_g>\" header", a); return QSB_ERR_PARSE; }
        (void)b; /* the gate count is implied by the text */
        c.nq = (int)a; s = e2;
    }
    int rc = parse_block(&c, s, text + strlen(text), NULL, NULL, 0);
    free(c.defs);
    if (rc) { free(c.gv.g); return rc; }
    if (c.nq < 0) { qsb_set_error("no qubit declaration found"); free(c.gv.g); return QSB_ERR_PARSE; }
    if (c.max_q >= c.nq) { qsb_set_error("operand q[%d] exceeds the declared %d qubits", c.max_q, c.nq); free(c.gv.g); return QSB_ERR_PARSE; }
    *num_qubits = c.nq; *gates = c.gv.g; *n = c.gv.n;
    if (!c.gv.g) *gates = (qsb_gate_t *)calloc(1, sizeof(qsb_gate_t));
    return QSB_OK;
}

int qsb_parse_qasm_file(const char *path, int *num_qubits, qsb_gate_t **gates, size_t *n)
{
    if (!path) { qsb_set_error("qsb_parse_qasm_file: null path"); return QSB_ERR_ARG; }
    FILE *f = fopen(path, "rb");
    if (!f) { qsb_set_error("ERROR: cannot open circuit file"); return QSB_ERR_IO; }
    /* growing read loop: works for pipes / FIFOs / /dev/stdin as well (ftell() is -1 there), checks read errors */
    size_t cap = 1 << 16, got = 0;
    char *buf = (char *)malloc(cap + 1);
    if (!buf) { fclose(f); qsb_set_error("Malloc error"); return QSB_ERR_NOMEM; }
    for (;;) {
        const size_t r = fread(buf + got, 1, cap - got, f);
        got += r;
        if (r == 0) {
            if (ferror(f)) { fclose(f); free(buf); qsb_set_error("ERROR: cannot read circuit file"); return QSB_ERR_IO; }
            break;
        }
        if (got == cap) {
            if (cap > ((size_t)1 << 40)) { fclose(f); free(buf); qsb_set_error("circuit file too large"); return QSB_ERR_NOMEM; }
            char *nb = (char *)realloc(buf, cap * 2 + 1);
            if (!nb) { fclose(f); free(buf); qsb_set_error("Malloc error"); return QSB_ERR_NOMEM; }
            buf = nb; cap *= 2;
        }
    }
    fclose(f);
    buf[got] = 0;
    if (memchr(buf, 0, got)) { free(buf); qsb_set_error("circuit file contains a NUL byte"); return QSB_ERR_PARSE; }
    int rc = qsb_parse_qasm_string(buf, num_qubits, gates, n);
    free(buf);
    return rc;
}
