// TEST INFRASTRUCTURE ONLY — stand-in for Xilinx <ap_fixed.h>; only the names
// that hlslib/include/hlslib/xilinx/DataPack.h:66-104 mentions in (never
// instantiated) partial specialisations are declared.
#pragma once
#include "ap_int.h"

enum ap_q_mode { AP_RND, AP_RND_ZERO, AP_RND_MIN_INF, AP_RND_INF, AP_RND_CONV, AP_TRN, AP_TRN_ZERO };
enum ap_o_mode { AP_SAT, AP_SAT_ZERO, AP_SAT_SYM, AP_WRAP, AP_WRAP_SM };

template <int W, int I, ap_q_mode Q = AP_TRN, ap_o_mode O = AP_WRAP, int N = 0>
struct ap_fixed;
template <int W, int I, ap_q_mode Q = AP_TRN, ap_o_mode O = AP_WRAP, int N = 0>
struct ap_ufixed;
