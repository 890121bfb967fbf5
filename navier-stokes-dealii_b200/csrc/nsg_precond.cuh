// nsg_precond.cuh — K7: ILU(0) + block preconditioners (src/NavierStokesSolver.hpp:520-639).
#pragma once
#include "nsg_common.cuh"

namespace nsg {
static int build_blocks(nsg_ctx *) { return NSG_OK; }
static void free_blocks(nsg_ctx *) {}
static int precond_initialize(nsg_ctx *) { return fail(NSG_ERR_STATE, "block preconditioners not built yet"); }
static int precond_vmult(nsg_ctx *, int, double *, const double *) { return fail(NSG_ERR_STATE, "block preconditioners not built yet"); }
static int ilu_apply(nsg_ctx *, CsrBlock &, double *, const double *) { return fail(NSG_ERR_STATE, "block preconditioners not built yet"); }
}  // namespace nsg
