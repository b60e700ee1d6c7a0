// hh_policy_tc.h -- internal interface between hh_policy.cu (C ABI, argument checks) and hh_policy_tc.cu (tcgen05 path)
#pragma once
#include <string>

#include "../../include/hhmarl_b200.h"

int hh_pf_tc_pack(const float* w_dev, int k_rows, int n_cols, int ldw, int n_total, int n_chunk, int row_shift, int ksteps, int kps,
                  void* image_dev, float* unscale_dev, void* stream, std::string& err);
int hh_pf_tc_launch(const hh_policy_chain_ex* chains, int n_chains, int max_rows, void* stream, std::string& err);
