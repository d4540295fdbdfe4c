/* Minimal hand-written declarations for the ABI of the system libgmp.so.10 (GMP 6.x).
 *
 * This image ships the GMP runtime (libgmp.so.10) but not its development header.
 * Both the reference build under oracle/_ref and the host-side drop-in layer
 * (c_lwe_snarks_b200/host) only need the public mpz_t layout and a few dozen entry
 * points, so they are declared here.  Link with `-l:libgmp.so.10`.
 *
 * When a real <gmp.h> is installed it can be used instead: every name below has the
 * same meaning there.
 */
#ifndef MFB200_COMPAT_GMP_H
#define MFB200_COMPAT_GMP_H

#include <stddef.h>
#include <stdint.h>
#include <stdarg.h>

#ifdef __cplusplus
extern "C" {
#endif

#define __GNU_MP_VERSION 6
#define GMP_LIMB_BITS 64
#define GMP_NUMB_BITS 64
#define GMP_NAIL_BITS 0

typedef unsigned long mp_limb_t;
typedef long mp_limb_signed_t;
typedef unsigned long mp_bitcnt_t;
typedef long mp_size_t;
typedef mp_limb_t *mp_ptr;
typedef const mp_limb_t *mp_srcptr;

typedef struct {
  int _mp_alloc; /* limbs allocated at _mp_d */
  int _mp_size;  /* abs() = limbs in use, sign = sign of the number */
  mp_limb_t *_mp_d;
} __mpz_struct;

typedef __mpz_struct mpz_t[1];
typedef __mpz_struct *mpz_ptr;
typedef const __mpz_struct *mpz_srcptr;

#define __GMP_LIKELY(cond) __builtin_expect((cond) != 0, 1)
#define __GMP_UNLIKELY(cond) __builtin_expect((cond) != 0, 0)

void __gmp_set_memory_functions(void *(*)(size_t), void *(*)(void *, size_t, size_t),
                                void (*)(void *, size_t));
void __gmp_get_memory_functions(void *(**)(size_t), void *(**)(void *, size_t, size_t),
                                void (**)(void *, size_t));
#define mp_set_memory_functions __gmp_set_memory_functions
#define mp_get_memory_functions __gmp_get_memory_functions

void __gmpz_init(mpz_ptr);
void __gmpz_init2(mpz_ptr, mp_bitcnt_t);
void __gmpz_inits(mpz_ptr, ...);
void __gmpz_clear(mpz_ptr);
void __gmpz_clears(mpz_ptr, ...);
void *__gmpz_realloc(mpz_ptr, mp_size_t);
void __gmpz_realloc2(mpz_ptr, mp_bitcnt_t);
void __gmpz_set(mpz_ptr, mpz_srcptr);
void __gmpz_set_ui(mpz_ptr, unsigned long);
void __gmpz_init_set_ui(mpz_ptr, unsigned long);
unsigned long __gmpz_get_ui(mpz_srcptr);
void __gmpz_swap(mpz_ptr, mpz_ptr);
void __gmpz_add(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_add_ui(mpz_ptr, mpz_srcptr, unsigned long);
void __gmpz_sub(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_sub_ui(mpz_ptr, mpz_srcptr, unsigned long);
void __gmpz_mul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_mul_ui(mpz_ptr, mpz_srcptr, unsigned long);
void __gmpz_mul_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void __gmpz_addmul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_addmul_ui(mpz_ptr, mpz_srcptr, unsigned long);
void __gmpz_submul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_neg(mpz_ptr, mpz_srcptr);
unsigned long __gmpz_fdiv_r_ui(mpz_ptr, mpz_srcptr, unsigned long);
void __gmpz_fdiv_r_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
void __gmpz_fdiv_q_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
unsigned long __gmpz_cdiv_q_ui(mpz_ptr, mpz_srcptr, unsigned long);
int __gmpz_cmp(mpz_srcptr, mpz_srcptr);
int __gmpz_cmp_ui(mpz_srcptr, unsigned long);
int __gmpz_tstbit(mpz_srcptr, mp_bitcnt_t);
void __gmpz_setbit(mpz_ptr, mp_bitcnt_t);
void __gmpz_clrbit(mpz_ptr, mp_bitcnt_t);
void __gmpz_ui_pow_ui(mpz_ptr, unsigned long, unsigned long);
size_t __gmpz_sizeinbase(mpz_srcptr, int);
void *__gmpz_export(void *, size_t *, int, size_t, int, size_t, mpz_srcptr);
void __gmpz_import(mpz_ptr, size_t, int, size_t, int, size_t, const void *);

#define mpz_init __gmpz_init
#define mpz_init2 __gmpz_init2
#define mpz_inits __gmpz_inits
#define mpz_clear __gmpz_clear
#define mpz_clears __gmpz_clears
#define _mpz_realloc __gmpz_realloc
#define mpz_realloc2 __gmpz_realloc2
#define mpz_set __gmpz_set
#define mpz_set_ui __gmpz_set_ui
#define mpz_init_set_ui __gmpz_init_set_ui
#define mpz_get_ui __gmpz_get_ui
#define mpz_swap __gmpz_swap
#define mpz_add __gmpz_add
#define mpz_add_ui __gmpz_add_ui
#define mpz_sub __gmpz_sub
#define mpz_sub_ui __gmpz_sub_ui
#define mpz_mul __gmpz_mul
#define mpz_mul_ui __gmpz_mul_ui
#define mpz_mul_2exp __gmpz_mul_2exp
#define mpz_addmul __gmpz_addmul
#define mpz_addmul_ui __gmpz_addmul_ui
#define mpz_submul __gmpz_submul
#define mpz_neg __gmpz_neg
#define mpz_fdiv_r_ui __gmpz_fdiv_r_ui
#define mpz_mod_ui __gmpz_fdiv_r_ui
#define mpz_fdiv_r_2exp __gmpz_fdiv_r_2exp
#define mpz_fdiv_q_2exp __gmpz_fdiv_q_2exp
#define mpz_cdiv_q_ui __gmpz_cdiv_q_ui
#define mpz_cmp __gmpz_cmp
#define mpz_cmp_ui __gmpz_cmp_ui
#define mpz_tstbit __gmpz_tstbit
#define mpz_setbit __gmpz_setbit
#define mpz_clrbit __gmpz_clrbit
#define mpz_ui_pow_ui __gmpz_ui_pow_ui
#define mpz_sizeinbase __gmpz_sizeinbase
#define mpz_export __gmpz_export
#define mpz_import __gmpz_import
#define mpz_sgn(z) ((z)->_mp_size < 0 ? -1 : (z)->_mp_size > 0)

#ifdef __cplusplus
}
#endif
#endif /* MFB200_COMPAT_GMP_H */
