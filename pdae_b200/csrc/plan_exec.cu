// pdae_plan_* -- a launch plan at the C-ABI level (plan_create / destroy / run_step).
//
// A plan is an ordered list of recorded entry-point calls (name + argument values; the stream slot is patched at run time).
// `pdae_plan_run_step` replays the list on the caller's stream from native code: one foreign call per network evaluation /
// training pass instead of one per kernel (the Python engine's ctypes loop costs ~10 us per op; a training pass has ~700).
// All buffers are the caller's (device pointers recorded by value); the plan owns only its own host-side list.  Not thread-safe
// per plan (one plan per stream), no global mutable state.
#include <string.h>

#include <vector>

#include "common.cuh"

namespace {

struct PlanEntry {
  const char* name;
  void* fn;
  int (*call)(void*, const pdae_arg*);
  int nargs;
  const char* sig;
};

#include "plan_exec_table.inc"

struct PlanOp {
  const PlanEntry* e;
  int stream_slot;
  std::vector<pdae_arg> args;
};

const PlanEntry* find_entry(const char* name) {
  for (const PlanEntry& e : kPlanEntries)
    if (strcmp(e.name, name) == 0) return &e;
  return nullptr;
}

}  // namespace

struct pdae_plan {
  std::vector<PlanOp> ops;
};

using namespace pdae;

extern "C" int pdae_plan_create(pdae_plan** plan_out) {
  PDAE_REQUIRE(plan_out, "plan_create: null pointer");
  *plan_out = new pdae_plan();
  return PDAE_OK;
}

extern "C" void pdae_plan_destroy(pdae_plan* plan) { delete plan; }

extern "C" int pdae_plan_size(const pdae_plan* plan) { return plan ? (int)plan->ops.size() : 0; }

extern "C" int pdae_plan_add(pdae_plan* plan, const char* entry, const pdae_arg* args, int nargs, int stream_slot) {
  PDAE_REQUIRE(plan && entry && (args || nargs == 0), "plan_add: null pointer");
  const PlanEntry* e = find_entry(entry);
  PDAE_REQUIRE(e != nullptr, "plan_add: '%s' is not a recordable entry point", entry);
  PDAE_REQUIRE(nargs == e->nargs, "plan_add: %s takes %d arguments, got %d", entry, e->nargs, nargs);
  PDAE_REQUIRE(stream_slot >= -1 && stream_slot < nargs && (stream_slot < 0 || e->sig[stream_slot] == 'p'),
               "plan_add: %s: bad stream slot %d", entry, stream_slot);
  PlanOp op;
  op.e = e;
  op.stream_slot = stream_slot;
  op.args.assign(args, args + nargs);
  plan->ops.push_back(op);
  return PDAE_OK;
}

extern "C" int pdae_plan_run_step(pdae_plan* plan, pdae_stream_t stream) {
  PDAE_REQUIRE(plan, "plan_run_step: null plan");
  for (size_t i = 0; i < plan->ops.size(); ++i) {
    PlanOp& op = plan->ops[i];
    if (op.stream_slot >= 0) op.args[op.stream_slot].p = stream;
    const int rc = op.e->call(op.e->fn, op.args.data());
    if (rc != PDAE_OK) return rc;      // pdae_last_error() holds the failing entry point's message
  }
  return PDAE_OK;
}

extern "C" const char* pdae_plan_op_name(const pdae_plan* plan, int index) {
  if (!plan || index < 0 || index >= (int)plan->ops.size()) return nullptr;
  return plan->ops[index].e->name;
}
