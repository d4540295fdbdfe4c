/* TEST INFRASTRUCTURE — CPU restatement of the reference's hot path (see mf_oracle.c).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load liboracle.so, and only as the checker.  The product (c_lwe_snarks_b200) never does.
 *
 * Flat formats (shared with oracle/ref_shim.c so the two can be compared byte for byte):
 *   coordinate  = ORC_LIMBS (12) little-endian uint64 limbs  (736-bit reference width)
 *   ciphertext  = ORC_NC (1471) coordinates: a_0..a_1469, b
 *   secret key  = ORC_N (1470) coordinates
 */
#ifndef MF_ORACLE_H
#define MF_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#define ORC_N 1470
#define ORC_NC (ORC_N + 1)
#define ORC_LIMBS 12
#define ORC_LOGQ 736
#define ORC_LOGQ_EFF 704
#define ORC_P 0xfffffffbULL
#define ORC_CT_BYTES 92
#define ORC_CTR_CT (ORC_CT_BYTES * ORC_N) /* 135240 stream bytes per ciphertext */
#define ORC_NOISE_BYTES 69                /* (556+3)/8, entropy.c:32 */
#define ORC_SMUDGE_BYTES 80               /* 640/8 */

void orc_aes256_encrypt_block(const uint8_t key[32], const uint8_t in[16], uint8_t out[16]);
void orc_stream(const uint8_t seed[40], uint64_t offset, uint8_t *out, size_t nbytes);
void orc_urandomb(const uint8_t seed[40], uint64_t offset, size_t nbits, uint64_t *out_limbs,
                  int32_t *siz);
void orc_modq(const uint64_t *in, int nlimbs, uint64_t *out_limbs, int32_t *siz);

void orc_ct_import(const uint8_t seed[40], uint64_t offset, const uint8_t *b92, uint64_t *out);
void orc_ct_export(const uint64_t *ct_flat, uint8_t *b92);
void orc_ct_mul_ui(const uint64_t *a, uint64_t b, uint64_t *out);
void orc_ct_addmul_ui(uint64_t *rop, const uint64_t *a, uint64_t b);
void orc_ct_add(const uint64_t *a, const uint64_t *b, uint64_t *out);
void orc_eval_poly(const uint8_t seed[40], uint64_t offset, const uint8_t *c8,
                   const uint64_t *coeffs, size_t d, uint64_t *rop_flat);
int orc_ct_smudge(uint64_t *ct_flat, const uint8_t entropy[ORC_SMUDGE_BYTES + 1]);

void orc_key_gen(const uint8_t *entropy, uint64_t *sk_flat);
void orc_encrypt(const uint8_t seed[40], uint64_t offset, const uint64_t *sk_flat,
                 const uint64_t *m, size_t count, const uint8_t *entropy, uint8_t *out_b92,
                 uint64_t *out_ct_flat);
void orc_dotp(const uint64_t *a_flat, const uint64_t *b_flat, size_t len, uint64_t *out_limbs);
uint64_t orc_decrypt(const uint64_t *sk_flat, const uint64_t *ct_flat, int b_negative);

#endif
