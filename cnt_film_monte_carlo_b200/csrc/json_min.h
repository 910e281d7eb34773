// json_min.h -- a small JSON reader/writer for the reference's input.json schema.
//
// The reference parses its input with a vendored third-party header (lib/json.hpp, nlohmann-json 3.0.1), which this
// repository does not copy.  The engine only needs objects, arrays, strings, numbers, booleans and null.
#pragma once
#include <cctype>
#include <cstdlib>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace cntmc {
namespace json {

struct Value {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool                                          b = false;
  double                                        num = 0;
  std::string                                   str;
  std::vector<Value>                            arr;
  std::vector<std::pair<std::string, Value>>    obj;  // insertion order kept for dump()

  bool         is_object() const { return kind == Object; }
  bool         is_array() const { return kind == Array; }
  bool         is_number() const { return kind == Number; }
  bool         is_string() const { return kind == String; }
  bool         contains(const std::string& k) const { return find(k) != nullptr; }
  const Value* find(const std::string& k) const {
    if (kind != Object) return nullptr;
    for (const auto& kv : obj)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
  const Value& at(const std::string& k) const {
    const Value* v = find(k);
    if (!v) throw std::invalid_argument("json: missing key \"" + k + "\"");
    return *v;
  }
  const Value& at(size_t i) const {
    if (kind != Array || i >= arr.size()) throw std::invalid_argument("json: array index out of range");
    return arr[i];
  }
  double as_number() const {
    if (kind != Number) throw std::invalid_argument("json: value is not a number");
    return num;
  }
  bool as_bool() const {
    if (kind != Bool) throw std::invalid_argument("json: value is not a boolean");
    return b;
  }
  const std::string& as_string() const {
    if (kind != String) throw std::invalid_argument("json: value is not a string");
    return str;
  }
};

class Parser {
  const std::string& s;
  size_t             i = 0;

  [[noreturn]] void fail(const char* what) const {
    std::ostringstream m;
    m << "json: " << what << " at offset " << i;
    throw std::invalid_argument(m.str());
  }
  void ws() {
    while (i < s.size() && std::isspace((unsigned char)s[i])) ++i;
  }
  bool eat(char c) {
    ws();
    if (i < s.size() && s[i] == c) {
      ++i;
      return true;
    }
    return false;
  }
  std::string string_body() {
    std::string out;
    while (i < s.size() && s[i] != '"') {
      char c = s[i++];
      if (c == '\\') {
        if (i >= s.size()) fail("bad escape");
        char e = s[i++];
        switch (e) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            if (i + 4 > s.size()) fail("bad \\u escape");
            unsigned cp = (unsigned)std::strtoul(s.substr(i, 4).c_str(), nullptr, 16);
            i += 4;
            if (cp < 0x80) {
              out += (char)cp;
            } else if (cp < 0x800) {
              out += (char)(0xC0 | (cp >> 6));
              out += (char)(0x80 | (cp & 0x3F));
            } else {
              out += (char)(0xE0 | (cp >> 12));
              out += (char)(0x80 | ((cp >> 6) & 0x3F));
              out += (char)(0x80 | (cp & 0x3F));
            }
            break;
          }
          default: out += e;
        }
      } else {
        out += c;
      }
    }
    if (i >= s.size()) fail("unterminated string");
    ++i;
    return out;
  }

 public:
  explicit Parser(const std::string& text) : s(text) {}
  Value parse_document() {
    Value v = value();
    ws();
    if (i != s.size()) fail("trailing characters");
    return v;
  }
  Value value() {
    ws();
    if (i >= s.size()) fail("unexpected end");
    Value v;
    const char c = s[i];
    if (c == '{') {
      ++i;
      v.kind = Value::Object;
      if (eat('}')) return v;
      do {
        ws();
        if (i >= s.size() || s[i] != '"') fail("expected key");
        ++i;
        std::string k = string_body();
        if (!eat(':')) fail("expected ':'");
        v.obj.emplace_back(std::move(k), value());
      } while (eat(','));
      if (!eat('}')) fail("expected '}'");
    } else if (c == '[') {
      ++i;
      v.kind = Value::Array;
      if (eat(']')) return v;
      do {
        v.arr.push_back(value());
      } while (eat(','));
      if (!eat(']')) fail("expected ']'");
    } else if (c == '"') {
      ++i;
      v.kind = Value::String;
      v.str = string_body();
    } else if (s.compare(i, 4, "true") == 0) {
      i += 4;
      v.kind = Value::Bool;
      v.b = true;
    } else if (s.compare(i, 5, "false") == 0) {
      i += 5;
      v.kind = Value::Bool;
      v.b = false;
    } else if (s.compare(i, 4, "null") == 0) {
      i += 4;
    } else {
      char*       end = nullptr;
      const char* start = s.c_str() + i;
      v.num = std::strtod(start, &end);  // correctly rounded, like the reference's reader
      if (end == start) fail("unexpected character");
      i += (size_t)(end - start);
      v.kind = Value::Number;
    }
    return v;
  }
};

inline Value parse(const std::string& text) { return Parser(text).parse_document(); }

inline void dump(const Value& v, std::ostream& os, int indent = 4, int depth = 0) {
  const std::string pad((size_t)indent * (depth + 1), ' '), pad_close((size_t)indent * depth, ' ');
  switch (v.kind) {
    case Value::Null: os << "null"; break;
    case Value::Bool: os << (v.b ? "true" : "false"); break;
    case Value::Number: {
      std::ostringstream t;
      t.precision(17);
      t << v.num;
      os << t.str();
      break;
    }
    case Value::String: {
      os << '"';
      for (char c : v.str) {
        if (c == '"' || c == '\\') os << '\\';
        os << c;
      }
      os << '"';
      break;
    }
    case Value::Array:
      os << "[";
      for (size_t k = 0; k < v.arr.size(); ++k) {
        os << (k ? "," : "") << "\n" << pad;
        dump(v.arr[k], os, indent, depth + 1);
      }
      if (!v.arr.empty()) os << "\n" << pad_close;
      os << "]";
      break;
    case Value::Object:
      os << "{";
      for (size_t k = 0; k < v.obj.size(); ++k) {
        os << (k ? "," : "") << "\n" << pad << '"' << v.obj[k].first << "\": ";
        dump(v.obj[k].second, os, indent, depth + 1);
      }
      if (!v.obj.empty()) os << "\n" << pad_close;
      os << "}";
      break;
  }
}

}  // namespace json
}  // namespace cntmc
