// Table-driven exp / log / expm1 for the FP64-pipe kernels (component evaluation, statistic).
//
// libdevice's exp and log evaluate a degree-11 / degree-7 polynomial whose coefficients the
// compiler materialises with two UMOV / IMAD.MOV per coefficient inside the loop (12 % of the
// instructions the unfold and statistic kernels execute, profiles/r01i_stalls.md), and on a chip
// that runs these kernels AT ITS POWER CAP (tools/power_probe.py: 997 W, 1.81 GHz) every FP64
// operation is paid for twice, in issue slots and in clock.  Here
//   exp(x)  = 2^m * T[j] * (1 + r q(r)),  k = rint(64 x / ln 2) = 64 m + j,  |r| <= ln 2 / 128,
//             q of degree 5: 11 FP64 operations and one 8-byte table load instead of 19;
//   log(x)  = e ln 2 + L[i] + log1p(r),  i = top 7 mantissa bits,  r = m / c_i - 1, |r| <= 2^-8,
//             log1p of degree 7, no division: 13 FP64 operations instead of 21 + a reciprocal.
// Tables hold correctly rounded entries (generated with 60-digit decimal arithmetic,
// tools/make_fastmath_tables.py); L[i] is -ln of the ROUNDED 1 / c_i, split hi + lo, so the
// argument reduction is exact.  Measured against glibc over 2e7 arguments (tools/fastmath_check.cc,
// also a CPU test): exp <= 1.0 ulp, log <= 1.0 ulp outside [0.98, 1.02] (inside, and for
// non-finite, denormal or non-positive arguments, the functions hand over to libm / libdevice).
// The reference computes with XLA's exp / log (~1 ulp); parity is held at 1e-10.
#pragma once

#if defined(__CUDA_ARCH__) || defined(__CUDACC_RTC__)
#define ELISA_FM_TABLE static __device__ const
// polynomial coefficients live in the constant bank: a DFMA takes them as c[bank][offset]
// operands, where an immediate costs two UMOV / IMAD.MOV per use
#define ELISA_FM_COEF static __constant__ double
#define ELISA_FM_FN __device__ __forceinline__
#else
#include <cmath>
#include <cstdint>
#include <cstring>
#define ELISA_FM_TABLE static const
#define ELISA_FM_COEF static const double
#define ELISA_FM_FN static inline
#endif

namespace elisa {
namespace fm {

// 2^(j/64), j = 0..63, correctly rounded
ELISA_FM_TABLE double kExp2Tab[64] = {0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0, 0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0, 0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0, 0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0, 0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0, 0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0, 0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0, 0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0, 0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0, 0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0, 0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0, 0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0, 0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0, 0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0, 0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0, 0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};
// 1 / c_i rounded, c_i = 1 + (i + 1/2) / 128; and -ln of exactly that rounded value, hi + lo
ELISA_FM_TABLE double kLogInv[128] = {0x1.fe01fe01fe020p-1, 0x1.fa11caa01fa12p-1, 0x1.f6310aca0dbb5p-1, 0x1.f25f644230ab5p-1, 0x1.ee9c7f8458e02p-1, 0x1.eae807aba01ebp-1, 0x1.e741aa59750e4p-1, 0x1.e3a9179dc1a73p-1, 0x1.e01e01e01e01ep-1, 0x1.dca01dca01dcap-1, 0x1.d92f2231e7f8ap-1, 0x1.d5cac807572b2p-1, 0x1.d272ca3fc5b1ap-1, 0x1.cf26e5c44bfc6p-1, 0x1.cbe6d9601cbe7p-1, 0x1.c8b265afb8a42p-1, 0x1.c5894d10d4986p-1, 0x1.c26b5392ea01cp-1, 0x1.bf583ee868d8bp-1, 0x1.bc4fd65883e7bp-1, 0x1.b951e2b18ff23p-1, 0x1.b65e2e3beee05p-1, 0x1.b37484ad806cep-1, 0x1.b094b31d922a4p-1, 0x1.adbe87f94905ep-1, 0x1.aaf1d2f87ebfdp-1, 0x1.a82e65130e159p-1, 0x1.a574107688a4ap-1, 0x1.a2c2a87c51ca0p-1, 0x1.a01a01a01a01ap-1, 0x1.9d79f176b682dp-1, 0x1.9ae24ea5510dap-1, 0x1.9852f0d8ec0ffp-1, 0x1.95cbb0be377aep-1, 0x1.934c67f9b2ce6p-1, 0x1.90d4f120190d5p-1, 0x1.8e6527af1373fp-1, 0x1.8bfce8062ff3ap-1, 0x1.899c0f601899cp-1, 0x1.87427bcc092b9p-1, 0x1.84f00c2780614p-1, 0x1.82a4a0182a4a0p-1, 0x1.8060180601806p-1, 0x1.7e225515a4f1dp-1, 0x1.7beb3922e017cp-1, 0x1.79baa6bb6398bp-1, 0x1.77908119ac60dp-1, 0x1.756cac201756dp-1, 0x1.734f0c541fe8dp-1, 0x1.713786d9c7c09p-1, 0x1.6f26016f26017p-1, 0x1.6d1a62681c861p-1, 0x1.6b1490aa31a3dp-1, 0x1.691473a88d0c0p-1, 0x1.6719f3601671ap-1, 0x1.6524f853b4aa3p-1, 0x1.63356b88ac0dep-1, 0x1.614b36831ae94p-1, 0x1.5f66434292dfcp-1, 0x1.5d867c3ece2a5p-1, 0x1.5babcc647fa91p-1, 0x1.59d61f123ccaap-1, 0x1.5805601580560p-1, 0x1.56397ba7c52e2p-1, 0x1.54725e6bb82fep-1, 0x1.52aff56a8054bp-1, 0x1.50f22e111c4c5p-1, 0x1.4f38f62dd4c9bp-1, 0x1.4d843bedc2c4cp-1, 0x1.4bd3edda68fe1p-1, 0x1.4a27fad76014ap-1, 0x1.4880522014880p-1, 0x1.46dce34596066p-1, 0x1.453d9e2c776cap-1, 0x1.43a2730abee4dp-1, 0x1.420b5265e5951p-1, 0x1.40782d10e6566p-1, 0x1.3ee8f42a5af07p-1, 0x1.3d5d991aa75c6p-1, 0x1.3bd60d9232955p-1, 0x1.3a524387ac822p-1, 0x1.38d22d366088ep-1, 0x1.3755bd1c945eep-1, 0x1.35dce5f9f2af8p-1, 0x1.34679ace01346p-1, 0x1.32f5ced6a1dfap-1, 0x1.3187758e9ebb6p-1, 0x1.301c82ac40260p-1, 0x1.2eb4ea1fed14bp-1, 0x1.2d50a012d50a0p-1, 0x1.2bef98e5a3711p-1, 0x1.2a91c92f3c105p-1, 0x1.293725bb804a5p-1, 0x1.27dfa38a1ce4dp-1, 0x1.268b37cd60127p-1, 0x1.2539d7e9177b2p-1, 0x1.23eb79717605bp-1, 0x1.22a0122a0122ap-1, 0x1.21579804855e6p-1, 0x1.2012012012012p-1, 0x1.1ecf43c7fb84cp-1, 0x1.1d8f5672e4abdp-1, 0x1.1c522fc1ce059p-1, 0x1.1b17c67f2bae3p-1, 0x1.19e0119e0119ep-1, 0x1.18ab083902bdbp-1, 0x1.1778a191bd684p-1, 0x1.1648d50fc3201p-1, 0x1.151b9a3fdd5c9p-1, 0x1.13f0e8d344724p-1, 0x1.12c8b89edc0acp-1, 0x1.11a3019a74826p-1, 0x1.107fbbe011080p-1, 0x1.0f5edfab325a2p-1, 0x1.0e40655826011p-1, 0x1.0d24456359e3ap-1, 0x1.0c0a7868b4171p-1, 0x1.0af2f722eecb5p-1, 0x1.09ddba6af8360p-1, 0x1.08cabb37565e2p-1, 0x1.07b9f29b8eae2p-1, 0x1.06ab59c7912fbp-1, 0x1.059eea0727586p-1, 0x1.04949cc1664c5p-1, 0x1.038c6b78247fcp-1, 0x1.02864fc7729e9p-1, 0x1.0182436517a37p-1, 0x1.0080402010080p-1};
ELISA_FM_TABLE double kLogHi[128] = {0x1.ff00aa2b10ba0p-9, 0x1.7dc475f810a69p-7, 0x1.3cea44346a584p-6, 0x1.b9fc027af919ap-6, 0x1.1b0d98923d97fp-5, 0x1.58a5bafc8e4d3p-5, 0x1.95c830ec8e3f2p-5, 0x1.d276b8adb0b56p-5, 0x1.075983598e471p-4, 0x1.253f62f0a1417p-4, 0x1.42edcbea646eep-4, 0x1.60658a93750c4p-4, 0x1.7da766d7b12d0p-4, 0x1.9ab42462033aep-4, 0x1.b78c82bb0eda0p-4, 0x1.d4313d66cb35dp-4, 0x1.f0a30c01162a4p-4, 0x1.0671512ca596fp-3, 0x1.14785846742acp-3, 0x1.2266f190a5acdp-3, 0x1.303d718e47fd5p-3, 0x1.3dfc2b0ecc62ap-3, 0x1.4ba36f39a55e5p-3, 0x1.59338d9982085p-3, 0x1.66acd4272ad51p-3, 0x1.740f8f54037a3p-3, 0x1.815c0a14357e9p-3, 0x1.8e928de886d41p-3, 0x1.9bb362e7dfb85p-3, 0x1.a8becfc882f19p-3, 0x1.b5b519e8fb5a6p-3, 0x1.c2968558c18c2p-3, 0x1.cf6354e09c5ddp-3, 0x1.dc1bca0abec7bp-3, 0x1.e8c0252aa5a60p-3, 0x1.f550a564b7b37p-3, 0x1.00e6c45ad501dp-2, 0x1.071b85fcd590dp-2, 0x1.0d46b579ab74bp-2, 0x1.136870293a8b0p-2, 0x1.1980d2dd4236fp-2, 0x1.1f8ff9e48a2f3p-2, 0x1.2596010df763ap-2, 0x1.2b9303ab89d25p-2, 0x1.31871c9544185p-2, 0x1.3772662bfd85cp-2, 0x1.3d54fa5c1f710p-2, 0x1.432ef2a04e813p-2, 0x1.49006804009d0p-2, 0x1.4ec9732600269p-2, 0x1.548a2c3add263p-2, 0x1.5a42ab0f4cfe2p-2, 0x1.5ff3070a793d4p-2, 0x1.659b57303e1f2p-2, 0x1.6b3bb2235943dp-2, 0x1.70d42e2789236p-2, 0x1.7664e1239dbcfp-2, 0x1.7bede0a37afbfp-2, 0x1.816f41da0d495p-2, 0x1.86e919a330ba1p-2, 0x1.8c5b7c858b48bp-2, 0x1.91c67eb45a83ep-2, 0x1.972a341135159p-2, 0x1.9c86b02dc0862p-2, 0x1.a1dc064d5b995p-2, 0x1.a72a4966bd9e9p-2, 0x1.ac718c258b0e5p-2, 0x1.b1b1e0ebdfc5ap-2, 0x1.b6eb59d3cf35cp-2, 0x1.bc1e08b0dad0ap-2, 0x1.c149ff115f027p-2, 0x1.c66f4e3ff6ff9p-2, 0x1.cb8e0744d7acap-2, 0x1.d0a63ae721e64p-2, 0x1.d5b7f9ae2c684p-2, 0x1.dac353e2c5955p-2, 0x1.dfc859906d5b5p-2, 0x1.e4c71a8687704p-2, 0x1.e9bfa659861f5p-2, 0x1.eeb20c640ddf3p-2, 0x1.f39e5bc811e5dp-2, 0x1.f884a36fe9ec1p-2, 0x1.fd64f20f61571p-2, 0x1.011fab125ff8ap-1, 0x1.0389eefce633cp-1, 0x1.05f14bd26459cp-1, 0x1.0855c884b450ep-1, 0x1.0ab76bece14d2p-1, 0x1.0d163ccb9d6b8p-1, 0x1.0f7241c9b497dp-1, 0x1.11cb81787ccf8p-1, 0x1.1422025243d45p-1, 0x1.1675cababa60ep-1, 0x1.18c6e0ff5cf07p-1, 0x1.1b154b57da29ep-1, 0x1.1d610fe677003p-1, 0x1.1faa34b87094cp-1, 0x1.21f0bfc65beecp-1, 0x1.2434b6f483934p-1, 0x1.26762013430e0p-1, 0x1.28b500df60783p-1, 0x1.2af15f02640acp-1, 0x1.2d2b4012edc9dp-1, 0x1.2f62a99509546p-1, 0x1.3197a0fa7fe6ap-1, 0x1.33ca2ba328994p-1, 0x1.35fa4edd36ea0p-1, 0x1.38280fe58797fp-1, 0x1.3a5373e7ebdf9p-1, 0x1.3c7c7fff73206p-1, 0x1.3ea33936b2f5bp-1, 0x1.40c7a4880dceap-1, 0x1.42e9c6ddf80bfp-1, 0x1.4509a5133bb0ap-1, 0x1.472743f33aaadp-1, 0x1.4942a83a2fc07p-1, 0x1.4b5bd6956e273p-1, 0x1.4d72d3a39fd01p-1, 0x1.4f87a3f5026e9p-1, 0x1.519a4c0ba3446p-1, 0x1.53aad05b99b7cp-1, 0x1.55b9354b40bcep-1, 0x1.57c57f336f191p-1, 0x1.59cfb25fae87fp-1, 0x1.5bd7d30e71c73p-1, 0x1.5ddde57149923p-1, 0x1.5fe1edad18919p-1, 0x1.61e3efda46467p-1};
ELISA_FM_TABLE double kLogLo[128] = {0x1.2821ad5a6d357p-63, 0x1.74944bc161072p-61, -0x1.865ad48159d00p-61, -0x1.90ae69229dc86p-60, -0x1.74d7444dd6241p-59, -0x1.cab8569c56e40p-64, 0x1.eb41d00a417e9p-60, 0x1.078f14c95ff53p-59, 0x1.006d2999e22dcp-58, 0x1.1f6d34e01d981p-61, -0x1.511583653349bp-58, -0x1.f108b1d8436d3p-59, 0x1.a2240644d7da2p-59, -0x1.a099e1c184e8ep-59, -0x1.3ef0e61f9b03cp-58, 0x1.b90dd951d90fap-58, 0x1.8be64b8b7759bp-59, -0x1.2f39b81479b67p-58, 0x1.94409f1d3f83ap-60, -0x1.dab840e7f6177p-57, -0x1.b5ae71f658247p-57, 0x1.ba62b8c13f7f4p-57, -0x1.f767e433c98aap-57, 0x1.8d16eaaba9419p-57, -0x1.9201c9c3d5165p-59, 0x1.6d9bf9d57b326p-58, 0x1.141b7f8c5fa9ep-58, 0x1.2589eb96a6240p-59, -0x1.51439c1ff83e7p-58, -0x1.a8c37918c39ebp-58, -0x1.d5d8023e61e5fp-57, 0x1.6108e3ae024acp-60, 0x1.339a07d55b696p-57, 0x1.c698a33316dfbp-58, -0x1.dc074737f9135p-60, -0x1.13a09202fe73dp-57, -0x1.3b9568ff6feadp-57, 0x1.08b83fcbdef40p-57, 0x1.21f640e1e5ec9p-56, 0x1.86cc531dba494p-57, -0x1.02c2e4f1b2eb9p-56, -0x1.93fbf3418960dp-57, -0x1.9eed8ae0ebd3cp-59, -0x1.85ad7f614ab51p-58, -0x1.ea3598981366fp-57, 0x1.02a7589fba088p-57, 0x1.53668e578d9cdp-58, -0x1.83262e2b59206p-57, -0x1.bff0d07c5df6dp-59, -0x1.1aa87d977dc5ep-56, -0x1.58ce7bf1846eep-56, -0x1.c6bcb7dee9a3dp-56, -0x1.063077d7e37b7p-56, 0x1.db0af8efb83c7p-62, 0x1.957a93326784dp-56, 0x1.ee99bf7143954p-56, -0x1.d6d5d64f5daf8p-57, -0x1.6783cb9801a5bp-56, 0x1.76dc35fb48fe4p-56, -0x1.700c9d2029045p-56, 0x1.d754b0205fa6cp-56, 0x1.5e3ea3b96a3dfp-57, -0x1.5a3f62db48f27p-56, 0x1.7e81149622bdfp-56, 0x1.a0128698ba0b8p-56, 0x1.529dac69f61f1p-56, 0x1.682c7ade8dee3p-56, -0x1.0ee1a7dd74ea6p-58, 0x1.1524332cd95c4p-56, -0x1.385e3e3ea99a8p-58, 0x1.46868de7f39f6p-57, -0x1.82947258b6889p-58, 0x1.c5bbc32ef5aebp-56, 0x1.4acce112c40f2p-57, 0x1.4841807b53f96p-57, -0x1.abc65a3f2f204p-56, 0x1.51e1399f96398p-56, -0x1.34c36e0f052b9p-56, -0x1.de45038241ecfp-56, -0x1.81e47141b8404p-56, 0x1.200e221139873p-59, 0x1.618ae4f008400p-56, -0x1.b615859d5a349p-62, 0x1.4043750211778p-55, 0x1.8aae29a41ba4ap-59, 0x1.935b8ee4f9efep-58, 0x1.785826e49f318p-55, 0x1.02936cabac09ap-56, 0x1.6119595d0f3c3p-59, 0x1.ba8443b9db19dp-55, 0x1.dc70f563f9920p-56, 0x1.7e5e3b6a496ecp-55, -0x1.cb19c15477c8ep-56, -0x1.9a6baf4f4e637p-56, 0x1.2770a5c124ab5p-56, 0x1.d27563647963dp-56, 0x1.c42f71ef43276p-55, -0x1.c24f0c9187c92p-57, -0x1.bebb8cf0f6d11p-57, -0x1.86a95781c6727p-56, 0x1.813f3f4aaa9a3p-60, 0x1.ed8322925675ap-56, 0x1.9ae9d3664e355p-55, -0x1.7dcbcc6300133p-55, 0x1.f6348fb97128fp-57, 0x1.1c6ba66fd0910p-55, 0x1.727d468096436p-56, -0x1.756f4d8a9b974p-57, 0x1.5ce11148e1124p-56, -0x1.e80db7025bed1p-60, 0x1.f66e975ec9f52p-59, 0x1.13c8b79ff2789p-58, -0x1.4d411c2cd7cf1p-55, -0x1.5701d7ad284a5p-55, -0x1.a930fed5d6b7ep-60, 0x1.2a18a88ca56b5p-56, -0x1.2c7a06beea772p-55, 0x1.01a9a829c011bp-56, -0x1.68ca8b1bcea9dp-55, 0x1.a332128e4a77fp-55, -0x1.7722c14b894e2p-57, -0x1.1f342e541a63dp-59, 0x1.1eac5c4377e6ep-55, -0x1.bb94822ace357p-57, -0x1.c9649352e8e44p-67, 0x1.0fa37d75ef285p-59, 0x1.92e93de3ce483p-56, 0x1.7923604841473p-57};
#define ELISA_FM_INV_L64 0x1.71547652b82fep+6  /* 64 / ln 2 */
#define ELISA_FM_L64_HI 0x1.62e42fee00000p-7  /* ln 2 / 64, 32 significant bits */
#define ELISA_FM_L64_LO 0x1.a39ef35793c76p-39
#define ELISA_FM_LN2_HI 0x1.62e42fefa2000p-1  /* ln 2, 40 significant bits */
#define ELISA_FM_LN2_LO 0x1.9ef35793c7673p-41

// 0-2: 64 / ln 2, -(ln 2 / 64) hi, lo; 3-7: 1/720 .. 1/2 (exp); 8-9: ln 2 hi, lo; 10-15: 1/7, -1/6, 1/5, -1/4, 1/3, -1/2 (log1p)
ELISA_FM_COEF kC[16] = {ELISA_FM_INV_L64, -ELISA_FM_L64_HI, -ELISA_FM_L64_LO, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5,
                        ELISA_FM_LN2_HI, ELISA_FM_LN2_LO, 1.0 / 7.0, -1.0 / 6.0, 0.2, -0.25, 1.0 / 3.0, -0.5};

#if defined(__CUDA_ARCH__) || defined(__CUDACC_RTC__)
ELISA_FM_FN int hi32(double x) { return __double2hiint(x); }
ELISA_FM_FN int lo32(double x) { return __double2loint(x); }
ELISA_FM_FN double mk(int hi, int lo) { return __hiloint2double(hi, lo); }
ELISA_FM_FN double fmad(double a, double b, double c) { return fma(a, b, c); }
ELISA_FM_FN double slow_exp(double x) { return ::exp(x); }
ELISA_FM_FN double slow_log(double x) { return ::log(x); }
ELISA_FM_FN double slow_expm1(double x) { return ::expm1(x); }
// 1 / a for a NORMAL a well inside the exponent range (the caller knows, e.g. clipped model counts or
// expm1(x) for x in (1e-4, 50)): MUFU.RCP64H seed (2^-23) and two Newton steps, <= 1 ulp, without the
// range checks and the final correction of the IEEE division -- 5 instructions where a quotient takes ~20
ELISA_FM_FN double rcp(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    return fma(r, e, r);
}
#else
ELISA_FM_FN int hi32(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
ELISA_FM_FN int lo32(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(b & 0xffffffff); }
ELISA_FM_FN double mk(int hi, int lo) { int64_t b = ((int64_t)hi << 32) | (uint32_t)lo; double x; std::memcpy(&x, &b, 8); return x; }
ELISA_FM_FN double fmad(double a, double b, double c) { return std::fma(a, b, c); }
ELISA_FM_FN double slow_exp(double x) { return std::exp(x); }
ELISA_FM_FN double slow_log(double x) { return std::log(x); }
ELISA_FM_FN double slow_expm1(double x) { return std::expm1(x); }
ELISA_FM_FN double rcp(double a) { return 1.0 / a; }
#endif

// exp(x).  Fast path for |x| < 700 (results in the normal range with room to spare).
ELISA_FM_FN double exp(double x) {
    if (!(fabs(x) < 700.0)) return slow_exp(x);  // also NaN
    const double SHIFT = 6755399441055744.0;     // 1.5 * 2^52: rint by addition
    const double t = fmad(x, kC[0], SHIFT);
    const int k = lo32(t);
    const double kf = t - SHIFT;
    double r = fmad(kf, kC[1], x);
    r = fmad(kf, kC[2], r);
    const double T = kExp2Tab[k & 63];
    double q = fmad(r, kC[3], kC[4]);
    q = fmad(q, r, kC[5]);
    q = fmad(q, r, kC[6]);
    q = fmad(q, r, kC[7]);
    q = fmad(q, r, 1.0);
    const double y = fmad(T, r * q, T);  // in [0.99, 2.02)
    return mk(hi32(y) + ((k >> 6) << 20), lo32(y));
}

// expm1(x): exp(x) - 1 loses nothing for |x| >= 0.25 (e^x - 1 is then at least 0.22 in magnitude
// against an error of half an ulp of e^x <= 1.3); below that the series / libm.
ELISA_FM_FN double expm1(double x) {
    if (!(fabs(x) >= 0.25 && fabs(x) < 700.0)) return slow_expm1(x);
    return exp(x) - 1.0;
}

// log(x) for normal positive finite x outside [0.98, 1.02]; libm otherwise.
ELISA_FM_FN double log(double x) {
    const int hi = hi32(x);
    // hi in [0x00100000, 0x7ff00000): positive, normal, finite
    if (!((unsigned)(hi - 0x00100000) < 0x7fe00000u) || (x > 0.98 && x < 1.02)) return slow_log(x);
    const int e = (hi >> 20) - 1023;
    const int i = (hi >> 13) & 127;
    const double m = mk((hi & 0x000fffff) | 0x3ff00000, lo32(x));  // [1, 2)
    const double r = fmad(m, kLogInv[i], -1.0);
    // exact int -> double without the quarter-rate I2F.F64: bits of 2^52 + (e + 2^31)
    const double ef = mk(0x43300000, e ^ (int)0x80000000) - 4503601774854144.0;  // 2^52 + 2^31
    double p = fmad(r, kC[10], kC[11]);
    p = fmad(p, r, kC[12]);
    p = fmad(p, r, kC[13]);
    p = fmad(p, r, kC[14]);
    p = fmad(p, r, kC[15]);
    const double r2 = r * r;
    const double hi_part = fmad(ef, kC[8], kLogHi[i]);  // exact product, one rounding
    const double lo_part = fmad(ef, kC[9], kLogLo[i]);
    return hi_part + (r + fmad(p, r2, lo_part));
}

}  // namespace fm
}  // namespace elisa
