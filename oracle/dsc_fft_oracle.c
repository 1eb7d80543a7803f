/* dsc_fft_oracle.c -- CPU restatement of DSC's FFT path in plain C.
 * TEST INFRASTRUCTURE ONLY; see dsc_fft_oracle.h for the rules and the parity status.
 * Citations are relative to /root/reference/.
 */
#include "dsc_fft_oracle.h"

#include <math.h>
#include <stdlib.h>

/* ---- integer rules ------------------------------------------------------------------ */

int dsco_pow2_n(int n)
{
    /* dsc.h:122-132 smears the top set bit of n-1 downwards and adds one. */
    if (n <= 0) return -1;
    unsigned v = (unsigned)(n - 1);
    for (int s = 1; s <= 16; s *= 2) v |= v >> s;
    return (int)(v + 1u);
}

int dsco_axis_index(int n_dim, int axis)
{
    /* dsc.h:81 -- shapes are right-aligned in a 4-slot array. */
    return axis < 0 ? 4 + axis : 4 - n_dim + axis;
}

int dsco_fft_len(int x_n, int n_arg)
{
    return n_arg > 0 ? dsco_pow2_n(n_arg) : dsco_pow2_n(x_n);
}

int dsco_rfft_len(int x_n, int n_arg, int *fft_order, int *out_n)
{
    const int order = dsco_pow2_n(n_arg > 0 ? n_arg : x_n) >> 1;
    if (order <= 0) return -1; /* dsc_plan_fft -> dsc_pow2_n(0) asserts, dsc.cpp:221 */
    *fft_order = order;
    *out_n = order + 1;
    return 0;
}

int dsco_irfft_len(int x_n, int n_arg, int *fft_order, int *out_n)
{
    const int bins = n_arg > 0 ? n_arg : x_n;
    if (bins - 1 <= 0) return -1; /* dsc_pow2_n(0) asserts */
    const int order = dsco_pow2_n(bins - 1);
    *fft_order = order;
    *out_n = order << 1;
    return 0;
}

size_t dsco_twiddle_count(int n, int real_plan)
{
    /* dsc_fft.h:116-121: sum of t for t = 2, 4, .., sets  ==  2*sets - 2 */
    size_t total = 0;
    const int sets = real_plan ? 2 * n : n;
    for (int t = 2; t <= sets; t *= 2) total += (size_t)t;
    return total;
}

/* ---- precision-generic bodies --------------------------------------------------------- */

#define R float
#define FN(name) name##_f32
#define R_COS cosf
#define R_SIN sinf
#define R_PI 3.14159265358979323846f
#include "dsc_fft_oracle_impl.inc"
#undef R
#undef FN
#undef R_COS
#undef R_SIN
#undef R_PI

#define R double
#define FN(name) name##_f64
#define R_COS cos
#define R_SIN sin
#define R_PI 3.14159265358979323846
#include "dsc_fft_oracle_impl.inc"
#undef R
#undef FN
#undef R_COS
#undef R_SIN
#undef R_PI

/* ---- dtype dispatch (dsc.cpp:2034-2068, 2207-2241) ------------------------------------- */

int dsco_fft(const void *x, int x_dtype, void *out,
             long outer, int x_n, long inner, int n_arg, int forward)
{
    if (x_n <= 0) return -1;
    const int fft_n = dsco_fft_len(x_n, n_arg);
    switch (x_dtype) {
    case DSCO_F32: return fft_tensor_f32((const float *)x, 0, (float *)out, outer, x_n, inner, fft_n, forward);
    case DSCO_C32: return fft_tensor_f32((const float *)x, 1, (float *)out, outer, x_n, inner, fft_n, forward);
    case DSCO_F64: return fft_tensor_f64((const double *)x, 0, (double *)out, outer, x_n, inner, fft_n, forward);
    case DSCO_C64: return fft_tensor_f64((const double *)x, 1, (double *)out, outer, x_n, inner, fft_n, forward);
    default: return -1;
    }
}

int dsco_rfft(const void *x, int x_dtype, void *out,
              long outer, int x_n, long inner, int n_arg)
{
    int order, out_n;
    if (x_n <= 0 || dsco_rfft_len(x_n, n_arg, &order, &out_n) != 0) return -1;
    switch (x_dtype) {
    case DSCO_F32: return rfft_tensor_f32((const float *)x, (float *)out, outer, x_n, inner, order, out_n);
    case DSCO_F64: return rfft_tensor_f64((const double *)x, (double *)out, outer, x_n, inner, order, out_n);
    default: return -1; /* "RFFT input must be real", dsc.cpp:2211 */
    }
}

int dsco_irfft(const void *x, int x_dtype, void *out,
               long outer, int x_n, long inner, int n_arg)
{
    int order, out_n;
    if (x_n <= 0 || dsco_irfft_len(x_n, n_arg, &order, &out_n) != 0) return -1;
    switch (x_dtype) {
    case DSCO_C32: return irfft_tensor_f32((const float *)x, (float *)out, outer, x_n, inner, order, out_n);
    case DSCO_C64: return irfft_tensor_f64((const double *)x, (double *)out, outer, x_n, inner, order, out_n);
    default: return -1; /* "IRFFT input must be complex", dsc.cpp:2215 */
    }
}
