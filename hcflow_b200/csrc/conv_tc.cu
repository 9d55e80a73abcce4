// placeholder until the tcgen05 kernel lands (replaced in the next commit)
#include "common.cuh"
struct hcf_conv_tc_plan { int dummy; };
extern "C" int hcf_conv_tc_supported(const hcf_conv_args*) { return 0; }
extern "C" int64_t hcf_conv_tc_weight_bytes(int32_t, int32_t) { return 0; }
extern "C" int hcf_conv_tc_pack_weights(const float*, int32_t, int32_t, float*) { return HCF_ENOTSUP; }
extern "C" int hcf_conv_tc_plan_create(const hcf_conv_args*, const float*, int32_t, hcf_conv_tc_plan**) { return HCF_ENOTSUP; }
extern "C" int hcf_conv_tc_run(const hcf_conv_tc_plan*, void*) { return HCF_ENOTSUP; }
extern "C" void hcf_conv_tc_plan_destroy(hcf_conv_tc_plan*) {}
