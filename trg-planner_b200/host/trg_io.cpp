// On-disk graph format of the reference: TRG::saveGraph / TRG::loadPrebuiltGraph
//   cpp/trg_planner/core/trg_planner/src/graph/trg.cpp:130-177 / :66-128
// JSON {"nodes":[{"id","pos":[x,y,z],"state"}], "edges":[{"source","target","weight","dist"}]}.
// The reference uses nlohmann::json (dump(4): keys sorted, floats widened to double and printed
// shortest-round-trip); this file writes the same layout and reads any JSON with that schema, so
// graphs built here load in an unmodified reference and vice versa.
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <variant>

#include "trg.h"

namespace {

// ---- minimal JSON value + recursive-descent reader ------------------------------------------
struct JValue;
using JArray  = std::vector<JValue>;
using JObject = std::map<std::string, JValue>;
struct JValue {
  std::variant<std::nullptr_t, bool, double, std::string, JArray, JObject> v;
  const JObject& obj() const { return std::get<JObject>(v); }
  const JArray& arr() const { return std::get<JArray>(v); }
  double num() const {
    if (auto* b = std::get_if<bool>(&v)) return *b ? 1.0 : 0.0;
    return std::get<double>(v);
  }
  const JValue& at(const std::string& k) const {
    auto it = obj().find(k);
    if (it == obj().end()) throw std::runtime_error("json: missing key '" + k + "'");
    return it->second;
  }
};

class JReader {
 public:
  explicit JReader(const std::string& s) : s_(s) {}
  JValue parse() {
    JValue v = value();
    ws();
    if (i_ != s_.size()) throw std::runtime_error("json: trailing characters");
    return v;
  }

 private:
  void ws() { while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\n' || s_[i_] == '\t' || s_[i_] == '\r')) ++i_; }
  char peek() { ws(); if (i_ >= s_.size()) throw std::runtime_error("json: unexpected end"); return s_[i_]; }
  void expect(char c) { if (peek() != c) throw std::runtime_error(std::string("json: expected '") + c + "'"); ++i_; }
  JValue value() {
    const char c = peek();
    if (c == '{') return object();
    if (c == '[') return array();
    if (c == '"') return JValue{str()};
    if (s_.compare(i_, 4, "true") == 0) { i_ += 4; return JValue{true}; }
    if (s_.compare(i_, 5, "false") == 0) { i_ += 5; return JValue{false}; }
    if (s_.compare(i_, 4, "null") == 0) { i_ += 4; return JValue{nullptr}; }
    return number();
  }
  JValue object() {
    expect('{');
    JObject o;
    if (peek() == '}') { ++i_; return JValue{o}; }
    while (true) {
      std::string k = str();
      expect(':');
      o[k] = value();
      if (peek() == ',') { ++i_; continue; }
      expect('}');
      break;
    }
    return JValue{std::move(o)};
  }
  JValue array() {
    expect('[');
    JArray a;
    if (peek() == ']') { ++i_; return JValue{a}; }
    while (true) {
      a.push_back(value());
      if (peek() == ',') { ++i_; continue; }
      expect(']');
      break;
    }
    return JValue{std::move(a)};
  }
  std::string str() {
    expect('"');
    std::string out;
    while (i_ < s_.size() && s_[i_] != '"') {
      if (s_[i_] == '\\' && i_ + 1 < s_.size()) {
        const char e = s_[i_ + 1];
        out += (e == 'n' ? '\n' : e == 't' ? '\t' : e);
        i_ += 2;
      } else {
        out += s_[i_++];
      }
    }
    if (i_ >= s_.size()) throw std::runtime_error("json: unterminated string");
    ++i_;
    return out;
  }
  JValue number() {
    ws();
    const char* b = s_.c_str() + i_;
    char* e = nullptr;
    const double d = std::strtod(b, &e);
    if (e == b) throw std::runtime_error("json: bad number");
    i_ += (size_t)(e - b);
    return JValue{d};
  }
  const std::string& s_;
  size_t i_ = 0;
};

}  // namespace

// nlohmann stores a float as double and prints a digit string that round-trips the double, laid out by its
// format_buffer (json.hpp, detail::to_chars): fixed notation while the decimal point lies within [-4, 15) digits of
// the first digit ("100000.0", "0.0001"), otherwise d[.ddd]e[+-]XX with at least two exponent digits. Same layout
// here; the digits are the SHORTEST round-trip string (std::to_chars), where nlohmann's Grisu2 now and then emits a
// 17th digit or rounds it the other way (2 % of random floats; both parse to the same double:
// tests/host/json_number_check.cpp).
std::string trg_b200::json_number(float f) {
  const double v = (double)f;
  if (v == 0.0) return std::signbit(v) ? "-0.0" : "0.0";
  if (!std::isfinite(v)) return "null";  // (nlohmann dumps non-finite numbers as null)
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof(buf), std::fabs(v), std::chars_format::scientific);  // d[.ddd]e[+-]XX, shortest
  std::string sci(buf, r.ptr);
  const size_t epos = sci.find('e');
  std::string digits;
  for (size_t i = 0; i < epos; ++i)
    if (sci[i] != '.') digits += sci[i];
  const int k = (int)digits.size();
  const int n = std::atoi(sci.c_str() + epos + 1) + 1;  // the decimal point sits after n digits
  std::string out = std::signbit(v) ? "-" : "";
  if (k <= n && n <= 15) {
    out += digits + std::string((size_t)(n - k), '0') + ".0";
  } else if (0 < n && n <= 15) {
    out += digits.substr(0, (size_t)n) + "." + digits.substr((size_t)n);
  } else if (-4 < n && n <= 0) {
    out += "0." + std::string((size_t)(-n), '0') + digits;
  } else {
    out += digits.substr(0, 1);
    if (k > 1) out += "." + digits.substr(1);
    const int e = n - 1;
    out += e < 0 ? "e-" : "e+";
    const int ae = e < 0 ? -e : e;
    if (ae < 10) out += "0";
    out += std::to_string(ae);
  }
  return out;
}

namespace {
inline std::string fnum(float f) { return trg_b200::json_number(f); }
}  // namespace

void TRG::saveGraph(const std::string& filepath) {  // trg.cpp:130-177
  std::lock_guard<std::mutex> lock(mtx.graph);
  std::filesystem::path save_path = filepath;
  if (save_path.extension().empty()) save_path += ".json";
  if (!save_path.parent_path().empty()) std::filesystem::create_directories(save_path.parent_path());
  trgStruct& g = *trgMap_["global"];
  std::ostringstream o;
  const char* I1 = "    ";
  const char* I2 = "        ";
  const char* I3 = "            ";
  const char* I4 = "                ";
  o << "{\n" << I1 << "\"edges\": [";
  bool first = true;
  for (const auto& kv : g.nodes) {
    const Node* n = kv.second;
    for (const Edge* e : n->edges_) {
      o << (first ? "\n" : ",\n") << I2 << "{\n"
        << I3 << "\"dist\": " << fnum(e->dist_) << ",\n"
        << I3 << "\"source\": " << n->id_ << ",\n"
        << I3 << "\"target\": " << e->dst_id_ << ",\n"
        << I3 << "\"weight\": " << fnum(e->weight_) << "\n"
        << I2 << "}";
      first = false;
    }
  }
  o << (first ? "],\n" : std::string("\n") + I1 + "],\n");
  o << I1 << "\"nodes\": [";
  first = true;
  for (const auto& kv : g.nodes) {
    const Node* n = kv.second;
    o << (first ? "\n" : ",\n") << I2 << "{\n"
      << I3 << "\"id\": " << n->id_ << ",\n"
      << I3 << "\"pos\": [\n"
      << I4 << fnum(n->pos_.x()) << ",\n"
      << I4 << fnum(n->pos_.y()) << ",\n"
      << I4 << fnum(n->pos_.z()) << "\n"
      << I3 << "],\n"
      << I3 << "\"state\": " << static_cast<int>(n->state_) << "\n"
      << I2 << "}";
    first = false;
  }
  o << (first ? "]\n" : std::string("\n") + I1 + "]\n") << "}";
  std::ofstream file(save_path);
  if (!file) throw std::runtime_error("trg_b200: cannot write " + save_path.string());
  file << o.str();
}

void TRG::loadPrebuiltGraph(const std::string& filepath) {  // trg.cpp:66-128
  std::lock_guard<std::mutex> lock(mtx.graph);
  std::filesystem::path load_path = filepath;
#ifdef TRG_DIR
  if (!std::filesystem::exists(load_path)) load_path = std::string(TRG_DIR) + "/../../" + filepath;  // trg.cpp:68
#endif
  if (!std::filesystem::exists(load_path)) throw std::runtime_error("trg_b200: File not found: " + load_path.string());
  std::ifstream file(load_path);
  std::stringstream ss;
  ss << file.rdbuf();
  const std::string text = ss.str();
  JValue root = JReader(text).parse();

  trgStruct& g = *trgMap_["global"];
  this->resetGraph("global");
  this->resetGraph("local");
  rewindPools();
  std::unordered_map<int, Node*> id_to_node;
  for (const JValue& nj : root.at("nodes").arr()) {
    const double id_d = nj.at("id").num();
    // ids index the CSR rows of the search graph (ensureDeviceGraph): a negative, fractional or repeated id
    // has no row; the reference would silently build a map it cannot plan on
    if (!(id_d >= 0.0 && id_d < 2147483647.0) || id_d != (double)(int)id_d)
      throw std::runtime_error("trg_b200: node id is not a non-negative 32-bit integer");
    const int id = (int)id_d;
    if (id_to_node.count(id)) throw std::runtime_error("trg_b200: node id appears twice");
    const JArray& p = nj.at("pos").arr();
    Eigen::Vector2f pos2d((float)p.at(0).num(), (float)p.at(1).num());
    NodeState state = static_cast<NodeState>((int)nj.at("state").num());
    Node* node  = newNode(id, pos2d, (float)p.at(2).num(), state);
    g.nodes[id] = node;
    id_to_node[id] = node;
    nodeIndexInsert(g, node);
    if (id >= g.node_id) g.node_id = id + 1;
  }
  for (const JValue& ej : root.at("edges").arr()) {
    const int src = (int)ej.at("source").num();
    const int dst = (int)ej.at("target").num();
    auto it = id_to_node.find(src);
    if (it == id_to_node.end()) throw std::runtime_error("trg_b200: edge source id not among nodes");
    if (!id_to_node.count(dst)) throw std::runtime_error("trg_b200: edge target id not among nodes");
    it->second->edges_.push_back(newEdge(dst, (float)ej.at("weight").num(), (float)ej.at("dist").num()));
  }
  invalidateDeviceGraph();
}
