// TEST INFRASTRUCTURE ONLY (oracle/_ref): the reference's logger, reduced to a sink.
#pragma once
#include <iostream>
#include <string>
namespace tlog {
struct Sink { template <typename T> Sink& operator<<(const T&) { return *this; } };
inline Sink info() { return {}; } inline Sink warning() { return {}; } inline Sink error() { return {}; } inline Sink success() { return {}; } inline Sink debug() { return {}; } inline Sink none() { return {}; }
}
