// fl_oracle_isa.cpp — TEST INFRASTRUCTURE (see fl_oracle_kernels.hpp header).
// One translation unit per (ISA level, element type): compiled with
//   -DFLO_NS=flo_v2|flo_v3|flo_v4  -DFLO_TBITS=8|16|32|64  -march=x86-64-v2|v3|v4
// and exports  flo_vK_run_uNN / flo_vK_single_uNN  for the dispatcher in fl_oracle.cpp.
#include "fl_oracle_kernels.hpp"

#define FLO_CAT_(a, b) a##b
#define FLO_CAT(a, b) FLO_CAT_(a, b)
#define FLO_CAT4_(a, b, c, d) a##b##c##d
#define FLO_CAT4(a, b, c, d) FLO_CAT4_(a, b, c, d)

#if FLO_TBITS == 8
using elem_t = uint8_t;
#elif FLO_TBITS == 16
using elem_t = uint16_t;
#elif FLO_TBITS == 32
using elem_t = uint32_t;
#elif FLO_TBITS == 64
using elem_t = uint64_t;
#else
#error "FLO_TBITS must be 8/16/32/64"
#endif

extern "C" {

void FLO_CAT4(FLO_NS, _run_u, FLO_TBITS, )(int op, unsigned width, size_t b0, size_t b1, const void* in,
                                            void* out, const void* base, const void* refs,
                                            uint64_t ref_scalar) {
    FLO_NS::run_blocks<elem_t>(op, width, b0, b1, static_cast<const elem_t*>(in), static_cast<elem_t*>(out),
                               static_cast<const elem_t*>(base), static_cast<const elem_t*>(refs),
                               elem_t(ref_scalar));
}

uint64_t FLO_CAT4(FLO_NS, _single_u, FLO_TBITS, )(unsigned width, const void* packed, size_t index) {
    return uint64_t(FLO_NS::unpack_single_rt<elem_t>(width, static_cast<const elem_t*>(packed), index));
}

}  // extern "C"
