/* bv_oracle.h -- TEST INFRASTRUCTURE ONLY (see bv_oracle.c). */
#ifndef BV_ORACLE_H
#define BV_ORACLE_H
#include <stdint.h>
#include "../include/basevar_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* One site from packed planes (same cell encoding as the device tiles). */
int bvo_site(const uint8_t* base, const uint8_t* qual, const uint8_t* strand, uint32_t n_samples,
             uint8_t ref_char, const bv_params* prm, bv_site_out* out);
/* A tile, site after site, single thread. */
int bvo_tile(const uint8_t* base, const uint8_t* qual, const uint8_t* strand, const uint8_t* ref_base,
             uint64_t pitch, uint32_t n_sites, uint32_t n_samples, const bv_params* prm, bv_site_out* out);
double bvo_gammaq(double s, double z);                                  /* kf_gammaq          */
double bvo_chi2_test(double chi, double dof);                           /* chi2_test          */
double bvo_fisher_two_sided(int n11, int n12, int n21, int n22);        /* fisher_exact_test  */
double bvo_fs_from_table(int ref_fwd, int ref_rev, int alt_fwd, int alt_rev);
double bvo_sor_from_table(int ref_fwd, int ref_rev, int alt_fwd, int alt_rev);
/* population group call: BaseType over the samples with sample_group[i] == g, lrt(order[0..n_order)) */
int bvo_group_site(const uint8_t* base, const uint8_t* qual, uint32_t n_samples, const uint8_t* sample_group,
                   uint32_t g, uint8_t ref_char, const uint8_t* order, int n_order, const bv_params* prm,
                   bv_group_out* out);
double bvo_erfc(double x);                                              /* kf_erfc            */
double bvo_wilcoxon(const double* s1, size_t n1, const double* s2, size_t n2);   /* wilcoxon_ranksum_test */
/* MQRankSum, ReadPosRankSum, BaseQRankSum of one called site (alt_mask bit b: base code b is a called ALT) */
int bvo_ranksums(const uint8_t* base, const uint8_t* qual, const uint8_t* mapq, const uint16_t* rpr, uint32_t n_samples,
                 uint8_t ref_char, uint32_t alt_mask, int32_t out3[3]);
#ifdef __cplusplus
}
#endif
#endif
