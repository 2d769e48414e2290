// TEST INFRASTRUCTURE ONLY (oracle/_ref): a small stand-in for nlohmann::json (absent dependency of the reference) — enough of its
// interface for the reference's own to_json / from_json functions (json_binding.h, cage.h, tet_mesh.h, affine_bounding_box.cuh) to run:
// a value tree (null, bool, number, string, array, object with insertion-ordered keys), conversion of C++ values through ADL-found
// to_json / from_json like nlohmann's adl_serializer, dump() / parse(). Written for this repository; not nlohmann's code.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace nlohmann {

class json {
public:
	enum class kind { null, boolean, number, string, array, object };

	json() = default;
	json(const json&) = default;
	json(json&&) = default;
	json& operator=(const json&) = default;
	json& operator=(json&&) = default;
	// any other value: through to_json(json&, const T&), found by argument-dependent lookup (nlohmann's adl_serializer)
	template <typename T, typename = typename std::enable_if<!std::is_same<typename std::decay<T>::type, json>::value>::type> json(const T& v) { assign(v); }
	template <typename T, typename = typename std::enable_if<!std::is_same<typename std::decay<T>::type, json>::value>::type> json& operator=(const T& v) {
		*this = json();
		assign(v);
		return *this;
	}

	static json array() { json j; j.k_ = kind::array; return j; }
	static json object() { json j; j.k_ = kind::object; return j; }

	bool is_null() const { return k_ == kind::null; }
	bool is_array() const { return k_ == kind::array; }
	bool is_object() const { return k_ == kind::object; }
	bool is_number() const { return k_ == kind::number; }
	bool is_string() const { return k_ == kind::string; }
	bool is_boolean() const { return k_ == kind::boolean; }
	size_t size() const { return k_ == kind::array ? arr_.size() : k_ == kind::object ? obj_.size() : (k_ == kind::null ? 0 : 1); }

	// object access: operator[] creates, at() throws
	json& operator[](const std::string& key) {
		if (k_ == kind::null) k_ = kind::object;
		if (k_ != kind::object) throw std::runtime_error("json: operator[](key) on a non-object");
		for (auto& kv : obj_) if (kv.first == key) return kv.second;
		obj_.emplace_back(key, json());
		return obj_.back().second;
	}
	json& operator[](const char* key) { return (*this)[std::string(key)]; }
	const json& operator[](const std::string& key) const { return at(key); }
	const json& operator[](const char* key) const { return at(std::string(key)); }
	const json& at(const std::string& key) const {
		if (k_ == kind::object) for (const auto& kv : obj_) if (kv.first == key) return kv.second;
		throw std::out_of_range("json: key '" + key + "' not found");
	}
	json& at(const std::string& key) { return const_cast<json&>(static_cast<const json&>(*this).at(key)); }
	const json& at(const char* key) const { return at(std::string(key)); }
	bool contains(const std::string& key) const {
		if (k_ == kind::object) for (const auto& kv : obj_) if (kv.first == key) return true;
		return false;
	}
	template <typename T> T value(const std::string& key, const T& d) const { return contains(key) ? at(key).template get<T>() : d; }

	// array access
	json& operator[](size_t i) { if (k_ != kind::array || i >= arr_.size()) throw std::out_of_range("json: index"); return arr_[i]; }
	const json& operator[](size_t i) const { return at(i); }
	json& operator[](int i) { return (*this)[(size_t)i]; }
	const json& operator[](int i) const { return at((size_t)i); }
	const json& at(size_t i) const { if (k_ != kind::array || i >= arr_.size()) throw std::out_of_range("json: index"); return arr_[i]; }
	const json& at(int i) const { return at((size_t)i); }
	void push_back(const json& v) {
		if (k_ == kind::null) k_ = kind::array;
		if (k_ != kind::array) throw std::runtime_error("json: push_back on a non-array");
		arr_.push_back(v);
	}
	template <typename T, typename = typename std::enable_if<!std::is_same<typename std::decay<T>::type, json>::value>::type> void push_back(const T& v) { push_back(json(v)); }
	const json* begin() const { return arr_.data(); }
	const json* end() const { return arr_.data() + arr_.size(); }

	// value out: through from_json(const json&, T&), found by ADL
	template <typename T> T get() const { T t{}; from_json(*this, t); return t; }
	template <typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value || std::is_same<T, std::string>::value>::type> operator T() const { return get<T>(); }
	template <typename T> void get_to(T& t) const { from_json(*this, t); }

	// raw accessors for the built-in conversions below
	kind type() const { return k_; }
	double number() const { if (k_ == kind::boolean) return b_ ? 1.0 : 0.0; if (k_ != kind::number) throw std::runtime_error("json: not a number"); return num_; }
	bool boolean() const { if (k_ == kind::number) return num_ != 0.0; if (k_ != kind::boolean) throw std::runtime_error("json: not a boolean"); return b_; }
	const std::string& str() const { if (k_ != kind::string) throw std::runtime_error("json: not a string"); return s_; }
	void set_number(double v) { *this = json(); k_ = kind::number; num_ = v; }
	void set_boolean(bool v) { *this = json(); k_ = kind::boolean; b_ = v; }
	void set_string(const std::string& v) { *this = json(); k_ = kind::string; s_ = v; }
	void set_array() { *this = json(); k_ = kind::array; }

	std::string dump(int /*indent*/ = -1) const { std::string out; write(out); return out; }
	static json parse(const std::string& text) {
		size_t p = 0;
		json j = parse_value(text, p);
		skip(text, p);
		if (p != text.size()) throw std::runtime_error("json: trailing characters");
		return j;
	}

private:
	template <typename T> void assign(const T& v) { to_json(*this, v); }

	void write(std::string& o) const {
		switch (k_) {
			case kind::null: o += "null"; break;
			case kind::boolean: o += b_ ? "true" : "false"; break;
			case kind::number: {
				char buf[40];
				if (num_ == (double)(long long)num_ && num_ > -1e15 && num_ < 1e15) snprintf(buf, sizeof buf, "%lld", (long long)num_);
				else snprintf(buf, sizeof buf, "%.17g", num_);
				o += buf;
				break;
			}
			case kind::string:
				o += '"';
				for (char c : s_) { if (c == '"' || c == '\\') { o += '\\'; o += c; } else if (c == '\n') o += "\\n"; else o += c; }
				o += '"';
				break;
			case kind::array:
				o += '[';
				for (size_t i = 0; i < arr_.size(); ++i) { if (i) o += ','; arr_[i].write(o); }
				o += ']';
				break;
			case kind::object:
				o += '{';
				for (size_t i = 0; i < obj_.size(); ++i) { if (i) o += ','; o += '"'; o += obj_[i].first; o += "\":"; obj_[i].second.write(o); }
				o += '}';
				break;
		}
	}
	static void skip(const std::string& t, size_t& p) { while (p < t.size() && (t[p] == ' ' || t[p] == '\n' || t[p] == '\t' || t[p] == '\r')) ++p; }
	static json parse_value(const std::string& t, size_t& p) {
		skip(t, p);
		if (p >= t.size()) throw std::runtime_error("json: unexpected end");
		json j;
		const char c = t[p];
		if (c == '{') {
			j.k_ = kind::object;
			++p; skip(t, p);
			if (t[p] == '}') { ++p; return j; }
			for (;;) {
				skip(t, p);
				json key = parse_value(t, p);
				skip(t, p);
				if (t[p] != ':') throw std::runtime_error("json: expected ':'");
				++p;
				j.obj_.emplace_back(key.str(), parse_value(t, p));
				skip(t, p);
				if (t[p] == ',') { ++p; continue; }
				if (t[p] == '}') { ++p; return j; }
				throw std::runtime_error("json: expected ',' or '}'");
			}
		}
		if (c == '[') {
			j.k_ = kind::array;
			++p; skip(t, p);
			if (t[p] == ']') { ++p; return j; }
			for (;;) {
				j.arr_.push_back(parse_value(t, p));
				skip(t, p);
				if (t[p] == ',') { ++p; continue; }
				if (t[p] == ']') { ++p; return j; }
				throw std::runtime_error("json: expected ',' or ']'");
			}
		}
		if (c == '"') {
			j.k_ = kind::string;
			++p;
			while (p < t.size() && t[p] != '"') {
				if (t[p] == '\\' && p + 1 < t.size()) { ++p; j.s_ += t[p] == 'n' ? '\n' : t[p]; }
				else j.s_ += t[p];
				++p;
			}
			++p;
			return j;
		}
		if (!t.compare(p, 4, "true")) { p += 4; j.k_ = kind::boolean; j.b_ = true; return j; }
		if (!t.compare(p, 5, "false")) { p += 5; j.k_ = kind::boolean; j.b_ = false; return j; }
		if (!t.compare(p, 4, "null")) { p += 4; return j; }
		char* end = nullptr;
		j.num_ = strtod(t.c_str() + p, &end);
		if (end == t.c_str() + p) throw std::runtime_error("json: bad token");
		p = (size_t)(end - t.c_str());
		j.k_ = kind::number;
		return j;
	}

	kind k_ = kind::null;
	bool b_ = false;
	double num_ = 0.0;
	std::string s_;
	std::vector<json> arr_;
	std::vector<std::pair<std::string, json>> obj_;
};

// ---- built-in conversions (nlohmann's own, in its namespace so that ADL on the json argument finds them) ----
template <typename T, typename std::enable_if<std::is_arithmetic<T>::value && !std::is_same<T, bool>::value, int>::type = 0> inline void to_json(json& j, const T& v) { j.set_number((double)v); }
inline void to_json(json& j, const bool& v) { j.set_boolean(v); }
inline void to_json(json& j, const std::string& v) { j.set_string(v); }
inline void to_json(json& j, const char* v) { j.set_string(v); }
template <typename T> inline void to_json(json& j, const std::vector<T>& v) {
	j.set_array();
	for (const T& e : v) { json x; to_json(x, e); j.push_back(x); }
}
template <typename T, typename std::enable_if<std::is_arithmetic<T>::value && !std::is_same<T, bool>::value, int>::type = 0> inline void from_json(const json& j, T& v) { v = (T)j.number(); }
inline void from_json(const json& j, bool& v) { v = j.boolean(); }
inline void from_json(const json& j, std::string& v) { v = j.str(); }
template <typename T> inline void from_json(const json& j, std::vector<T>& v) {
	v.clear();
	v.reserve(j.size());
	for (size_t i = 0; i < j.size(); ++i) { T e{}; from_json(j.at(i), e); v.push_back(std::move(e)); }
}

}  // namespace nlohmann
