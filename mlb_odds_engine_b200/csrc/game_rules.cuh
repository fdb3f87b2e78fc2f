// game_rules.cuh — base-running transition semantics shared by the native and replay kernels.
// Reference: core/half_inning_simulator.py:42-136 (handlers) — occupancy bitmask b0 = 1B, b1 = 2B, b2 = 3B.
#pragma once
#include <stdint.h>

#include "../../include/mlbsim.h"

// LUT entry: new_ob | runs << 16 | reached << 24 | three_outs << 30
__device__ inline uint32_t transition_entry(int o, int dA, int dB, int outs, int bases) {
  int nb = bases, r = 0, oa = 0;
  const int on1 = bases & 1, on2 = (bases >> 1) & 1, on3 = (bases >> 2) & 1;
  if (o == MLBSIM_K || o == MLBSIM_OUT) {  // half_inning_simulator.py:73-82
    if (on1 && outs < 2 && dA) { nb = bases & 6; oa = 2; } else { oa = 1; }
  } else if (o == MLBSIM_BB) {             // :84-94
    r = (bases == 7);
    nb = 1 | (on1 << 1) | ((on2 | on3) << 2);  // runner on 2B always goes to 3B; 3B holds unless forced
  } else if (o == MLBSIM_1B) {             // :97-109 (+ :28-39 folded into dA at two outs)
    r = on3 + (on2 & dA);
    nb = 1 | ((on1 & dB) << 1) | ((on2 & (dA ^ 1)) << 2);
  } else if (o == MLBSIM_2B) {             // :112-123
    r = on3 + on2 + (on1 & dA);
    nb = 2 | ((on1 & (dA ^ 1)) << 2);
  } else if (o == MLBSIM_3B) {             // :126-129
    r = on1 + on2 + on3; nb = 4;
  } else {                                 // HR :132-136
    r = on1 + on2 + on3 + 1; nb = 0;
  }
  int no = outs + oa;
  if (no > 3) no = 3;
  const int reached = (r > 0) || (nb != 0);
  return (uint32_t)(no * 8 + nb) | ((uint32_t)r << 16) | ((uint32_t)reached << 24) | ((uint32_t)(no == 3) << 30);
}

