/* qsb_internal.h -- declarations shared by the C and C++/CUDA parts of libqsim_b200. */
#ifndef QSB_INTERNAL_H
#define QSB_INTERNAL_H

#include "qsim_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

void qsb_set_error(const char *fmt, ...);

#ifdef __cplusplus
}
#endif
#endif
