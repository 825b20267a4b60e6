// oracle/ref_shim/g2o/types/sim3/types_seven_dof_expmap.h — TEST INFRASTRUCTURE ONLY: the two g2o types Utils/converter.h names in its
// declarations (nothing on the hot path converts through them).
#pragma once
namespace g2o {
class SE3Quat {};
class Sim3 {};
}  // namespace g2o
