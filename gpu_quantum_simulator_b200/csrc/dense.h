/* dense.h -- QSB_MODE_DENSE (experiment, see dense.cu): greedy dense k-qubit fusion + one sweep per block. */
#pragma once
#include <vector>
#include "common.cuh"

#define QSB_DENSE_MAX_K 5
struct DenseBlock {
    int k = 0;
    int q[QSB_DENSE_MAX_K] = {0, 0, 0, 0, 0};   /* logical qubits, ascending: bit j of the matrix index is qubit q[j] */
    std::vector<double> m;                      /* 2^k x 2^k row-major, (re, im) */
};
int dense_fuse(const std::vector<COp> &cops, const double gphase[2], int n, int k, std::vector<DenseBlock> &out);
struct qsb_sim;
int dense_execute(qsb_sim *s, const std::vector<DenseBlock> &blocks);
