// Shadows the reference's json_binding.h: TEST INFRASTRUCTURE ONLY (oracle/_ref). The real header pulls in the dataset loader (absent
// filesystem submodule); what the edits I/O needs from it — the Eigen <-> json glue (json_binding.h:27-75) and BoundingBox's (:79-87) — is cut
// out of the real file at build time (oracle/ref_build.py, json_binding.inc) and included here, over oracle/ref_shim/json/json.hpp.
#pragma once
#include <json/json.hpp>
#include <neural-graphics-primitives/common.h>
#include <neural-graphics-primitives/bounding_box.cuh>
#include "json_binding.inc"
