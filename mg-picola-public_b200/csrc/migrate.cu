// Slab ownership and particle migration (MoveParticles, auxPM.c:108-275).  Filled in below.
#include "common.cuh"

namespace mgp {

void particles_migrate(Ctx &c) {
  if (c.P == 1) return;
  throw Error(MGP_ERR_INVALID, "particle migration for nranks > 1 is not built yet");
}

}  // namespace mgp
