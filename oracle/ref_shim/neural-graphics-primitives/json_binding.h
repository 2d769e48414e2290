// Shadows the reference's json_binding.h (nlohmann glue for Eigen types): TEST INFRASTRUCTURE ONLY (oracle/_ref), inert.
#pragma once
#include <json/json.hpp>
#include <neural-graphics-primitives/common.h>
NGP_NAMESPACE_BEGIN
template <typename T> inline void to_json(nlohmann::json&, const T&) {}
template <typename T> inline void from_json(const nlohmann::json&, T&) {}
NGP_NAMESPACE_END
