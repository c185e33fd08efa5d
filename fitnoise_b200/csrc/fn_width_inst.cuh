// Instantiates the per-gene kernels for ONE design width (FN_W) and publishes them as a KernelSet.
// Included once per width: by fn_width.cu (one translation unit per width, the normal build) or,
// with -DFN_SINGLE_TU, eight times by fn_api.cu (one-command builds, e.g. the FN_TRACE build).
#ifndef FN_W
#error "define FN_W (design width) before including fn_width_inst.cuh"
#endif
#define FN_WCAT2(a, b) a##b
#define FN_WCAT(a, b) FN_WCAT2(a, b)

namespace fn {
const KernelSet* FN_WCAT(kernel_set_w, FN_W)() {
    static const KernelSet ks = {
        (const void*)prepare_kernel<FN_W>,
        {(const void*)eval_kernel<FN_W, KIND_SCALE1>, (const void*)eval_kernel<FN_W, KIND_SCALE_PS>,
         (const void*)eval_kernel<FN_W, KIND_PATSEQ>,
#if FN_W <= 4
         (const void*)eval_kernel<FN_W, KIND_PATSEQ_SHARED>
#else
         nullptr
#endif
        },
#if FN_W <= 8
        {(const void*)eval_kernel<FN_W, KIND_SCALE1, true>, (const void*)eval_kernel<FN_W, KIND_SCALE_PS, true>},
#else
        {nullptr, nullptr},
#endif
        (const void*)eval_factors_kernel<FN_W>,
        (const void*)coef_kernel<FN_W>,
    };
    return &ks;
}
}  // namespace fn
#undef FN_WCAT
#undef FN_WCAT2
