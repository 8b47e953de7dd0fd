/* mkl_cblas.h -- STAND-IN written for the oracle build (oracle/ref_shim_gemm_conv.cpp), NOT Intel's header.
 * The reference's x86 gemm-based INT8 convolution (saber/funcs/impl/x86/gemm_x8s8s32x_conv.cpp) calls exactly
 * one MKL routine, cblas_gemm_s8u8s32; MKL is not vendored in the reference and not installed here. The routine is
 * pure integer arithmetic with a documented definition (Intel MKL developer reference, "cblas_gemm_s8u8s32"):
 *     C := alpha * (op(A) + ao) * (op(B) + bo) + beta * C + C_offset
 * with A int8, B uint8, C int32 and C_offset a scalar (CblasFixOffset), a length-m column (CblasColOffset) or a
 * length-n row (CblasRowOffset). ref_shim_gemm_conv.cpp implements it naively. TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_STANDIN_MKL_CBLAS_H
#define ORACLE_STANDIN_MKL_CBLAS_H
#ifdef __cplusplus
extern "C" {
#endif
typedef int MKL_INT;
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_LAYOUT;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
typedef enum { CblasRowOffset = 171, CblasColOffset = 172, CblasFixOffset = 173 } CBLAS_OFFSET;
void cblas_gemm_s8u8s32(const CBLAS_LAYOUT layout, const CBLAS_TRANSPOSE transa, const CBLAS_TRANSPOSE transb,
                        const CBLAS_OFFSET offsetc, const MKL_INT m, const MKL_INT n, const MKL_INT k,
                        const float alpha, const void* a, const MKL_INT lda, const char ao, const void* b,
                        const MKL_INT ldb, const char bo, const float beta, int* c, const MKL_INT ldc, const int* co);
#ifdef __cplusplus
}
#endif
#endif
