// Parameters of the blocked Gauss-Seidel plans as the kernels need them (stage / window sizes of block_gs.cuh) with their
// environment overrides: shared by the upload path (engine.cu) and the host-only plan check (abi_plans.cu).
#pragma once
#include "engine_base.h"
#include "block_plan.h"
#include "pass_plan.h"

using namespace b200amg;

static BlockPlanParams block_params_from_env() {
  BlockPlanParams prm;
  prm.stage_nnz = kBgStageNnz; prm.stage_rows = kBgStageRows; prm.window = kBgWindow; prm.depth = kBgDepth;
  prm.step_us = 1e-3 * env_int("B200AMG_BLOCK_STEP_NS", 220);
  prm.cta_gbs = env_int("B200AMG_BLOCK_CTA_GBS", 55);
  prm.cap_step_to_stage = env_int("B200AMG_BLOCK_XCAP", 1);
  prm.force_tile_rows = env_int("B200AMG_BLOCK_TILE_ROWS", 0);
  prm.force_a = env_int("B200AMG_BLOCK_A", 0);
  prm.force_b = env_int("B200AMG_BLOCK_B", 0);
  prm.max_lanes = 32;
  prm.verbose = env_int("B200AMG_BLOCK_VERBOSE", 0);
  return prm;
}
