/* oracle_real.h -- shared scalar-type switch for the CPU oracle (test infrastructure, NOT product code). */
#ifndef ORACLE_REAL_H
#define ORACLE_REAL_H
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#ifdef ORACLE_F64
typedef double real;
#define R_SQRT sqrt
#define R_EXP exp
#define R_CEIL ceil
#define R_FLOOR floor
#define SUFFIX(name) name##_f64
#else
typedef float real;
#define R_SQRT sqrtf
#define R_EXP expf
#define R_CEIL ceilf
#define R_FLOOR floorf
#define SUFFIX(name) name##_f32
#endif
/* every literal is the fp32 literal of the spec, widened if real==double */
#define RC(x) ((real)(x##f))

#endif
