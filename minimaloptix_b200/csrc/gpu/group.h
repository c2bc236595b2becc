// group.h — several GPUs behind one handle (mox_create_multi, include/mox.h): the shell context
// returned to the caller owns one ordinary context per device and forwards every call.
#pragma once
#include <functional>
#include <string>
#include "mox.h"

struct mox_group;

int groupCreate(mox_group** out, const int* deviceIds, int n, std::string& err);
void groupDestroy(mox_group*);
int groupCount(const mox_group*);
mox_ctx* groupChild(mox_group*, int i);
const std::string& groupError(const mox_group*);

// f on every child, one after the other (host-side staging calls); stops at the first failure.
int groupEach(mox_group*, const std::function<int(mox_ctx*, int)>& f);
// f on every child at once, one persistent host thread per device (build, launch, pushes).
int groupParallel(mox_group*, const std::function<int(mox_ctx*, int)>& f);

int groupReadBegin(mox_group*);
int groupReadEnd(mox_group*, const float** out);
int groupStats(mox_group*, mox_stats* out);

// Internals of a plain context the group needs (defined in capi.cu).
extern "C" {
float* ctxGatherBuffer(mox_ctx*, int which);                // this context's own gather buffer (allocated on demand), or null
int ctxBorrowGatherTarget(mox_ctx*, int which, float* ptr); // push into a buffer owned by another context of this process
}
