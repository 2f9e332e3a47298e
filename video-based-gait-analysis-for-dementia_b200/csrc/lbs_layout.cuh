// Layout of the tensor-core skinning operand shared by its producer (smpl_pose_chain_kernel, smpl.cu) and its consumer
// (smpl_lbs_tc_kernel, lbs_tc.cu): per LBS_TC_FT frames one blob [hi | lo][k-chunk (6)][row group][8 rows][4 floats], row
// n = (f % LBS_TC_FT) * 12 + c (c = entry of the 3x4 skinning transform), k = joint: the UMMA no-swizzle K-major core-matrix order.
#pragma once
namespace gait {
constexpr int LBS_TC_FT = 16;                                   // frames per blob = per work item (UMMA N = 192)
constexpr int LBS_TC_AOP_PART_FLOATS = 6 * LBS_TC_FT * 12 * 4;     // one of the hi / lo halves
constexpr int LBS_TC_AOP_BLOB_BYTES = 2 * LBS_TC_AOP_PART_FLOATS * 4;
}  // namespace gait
