// TEST INFRASTRUCTURE ONLY (oracle/_ref): the reference's editing headers declare to_json/from_json members; nothing on the
// compiled path serialises, so nlohmann::json is an inert value here.
#pragma once
#include <string>
#include <vector>
namespace nlohmann {
struct json {
	template <typename T> json& operator=(const T&) { return *this; }
	json() = default;
	template <typename T> json(const T&) {}
	json& operator[](const std::string&) { return *this; }
	json& operator[](const char*) { return *this; }
	json& operator[](size_t) { return *this; }
	const json& operator[](const std::string&) const { return *this; }
	const json& operator[](const char*) const { return *this; }
	const json& at(const std::string&) const { return *this; }
	const json& at(size_t) const { return *this; }
	json& at(const std::string&) { return *this; }
	bool contains(const std::string&) const { return false; }
	template <typename T> T value(const std::string&, const T& d) const { return d; }
	template <typename T> T get() const { return T(); }
	template <typename T> operator T() const { return T(); }
	size_t size() const { return 0; }
	void push_back(const json&) {}
	bool is_null() const { return true; }
	static json array() { return {}; }
	static json object() { return {}; }
	const json* begin() const { return this; }
	const json* end() const { return this; }
};
}
