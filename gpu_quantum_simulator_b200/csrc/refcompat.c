/*
 * refcompat.c -- libqsim_b200_refcompat.so: the EXACT symbol names of
 * /root/reference/quantum_simulator.c:25-30, so that file's own main() (or any
 * caller written against it) links to the GPU path unchanged:
 *
 *     gcc -Dcompute_state_vector=unused_csv ...
 *
 * `complex` in the reference is `double _Complex`; the layouts match double[2].
 */
#include <complex.h>
#include <stdio.h>
#include <stdlib.h>

#include "qsim_b200.h"

double complex *compute_state_vector(char *filename, int *num_q)
{
    double *v = qsb_ref_compute_state_vector(filename, num_q);
    if (!v) {
        const char *e = qsb_last_error();
        printf("%s\n", e);
        if (e[0] == 'E') exit(1);      /* "ERROR: cannot open circuit file" exits in the reference (:129-130) */
    }
    return (double complex *)v;
}
void execute_single_qubit_gate(double complex *v, int num_q, double complex U[4], int target)
{
    qsb_ref_execute_single_qubit_gate((double *)v, num_q, (const double *)U, target);
}
void execute_cnot(double complex *v, int num_q, int control, int target)
{
    qsb_ref_execute_cnot((double *)v, num_q, control, target);
}
double *compute_state_cumulative_distribution(double complex *v, int num_q)
{
    return qsb_ref_compute_state_cumulative_distribution((const double *)v, num_q);
}
long long int measurement(double *cumul, int num_q) { return qsb_ref_measurement(cumul, num_q); }
