#include "map_io.h"

#include <cctype>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

namespace trg_b200 {

namespace {

std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a])) ++a;
  while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}

// "section.key" -> scalar text for a YAML file made of nested `key: value` maps (what config/*.yaml use)
std::map<std::string, std::string> flat_yaml(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("trg_b200: cannot open config " + path);
  std::map<std::string, std::string> out;
  std::vector<std::pair<int, std::string>> stack;  // (indent, key)
  std::string line;
  while (std::getline(f, line)) {
    {  // strip a comment: the first '#' that is not inside a quoted scalar
      char quote = 0;
      for (size_t i = 0; i < line.size(); ++i) {
        const char c = line[i];
        if (quote) { if (c == quote) quote = 0; }
        else if (c == '"' || c == '\'') quote = c;
        else if (c == '#') { line = line.substr(0, i); break; }
      }
    }
    if (trim(line).empty()) continue;
    int indent = 0;
    while (indent < (int)line.size() && line[indent] == ' ') ++indent;
    const size_t colon = line.find(':', indent);
    if (colon == std::string::npos) continue;
    const std::string key = trim(line.substr(indent, colon - indent));
    std::string val = trim(line.substr(colon + 1));
    while (!stack.empty() && stack.back().first >= indent) stack.pop_back();
    if (val.empty()) {
      stack.push_back({indent, key});
      continue;
    }
    if (val.size() >= 2 && (val.front() == '"' || val.front() == '\'') && val.back() == val.front()) val = val.substr(1, val.size() - 2);
    std::string full;
    for (auto& s : stack) full += s.second + ".";
    out[full + key] = val;
  }
  return out;
}

bool as_bool(const std::string& v) {
  std::string s;
  for (char c : v) s += (char)std::tolower((unsigned char)c);
  return s == "true" || s == "yes" || s == "on" || s == "y" || s == "1";
}

// LZF decompression (binary_compressed PCD payload), the format of liblzf / pcl::lzfDecompress
size_t lzf_decompress(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
  const uint8_t* ip = in;
  const uint8_t* const in_end = in + in_len;
  uint8_t* op = out;
  uint8_t* const out_end = out + out_len;
  while (ip < in_end) {
    unsigned ctrl = *ip++;
    if (ctrl < 32) {  // literal run of ctrl+1 bytes
      ++ctrl;
      if (op + ctrl > out_end || ip + ctrl > in_end) return 0;
      std::memcpy(op, ip, ctrl);
      op += ctrl;
      ip += ctrl;
    } else {  // back reference
      unsigned len = ctrl >> 5;
      if (ip >= in_end) return 0;
      if (len == 7) {
        len += *ip++;
        if (ip >= in_end) return 0;
      }
      const uint8_t* ref = op - ((ctrl & 0x1f) << 8) - 1 - *ip++;
      if (ref < out || op + len + 2 > out_end) return 0;
      len += 2;
      while (len--) *op++ = *ref++;
    }
  }
  return (size_t)(op - out);
}

}  // namespace

PlannerParams load_params_yaml(const std::string& config_path) {  // trg_planner.cpp:103-129
  const auto y = flat_yaml(config_path);
  PlannerParams p;
  auto has = [&](const char* k) { return y.find(k) != y.end(); };
  auto f = [&](const char* k, float d) { return has(k) ? std::stof(y.at(k)) : d; };
  auto b = [&](const char* k, bool d) { return has(k) ? as_bool(y.at(k)) : d; };
  auto s = [&](const char* k) { return has(k) ? y.at(k) : std::string(); };
  p.isVerbose     = b("isVerbose", true);
  p.graph_rate    = f("timer.graphRate", 1.0f);
  p.planning_rate = f("timer.planningRate", 1.0f);
  p.isPreMap      = b("map.isPrebuiltMap", false);
  p.preMapPath    = s("map.prebuiltMapPath");
  p.isVoxelize    = b("map.isVoxelize", false);
  p.VoxelSize     = f("map.voxelSize", 0.1f);
  p.isPreGraph    = b("trg.isPrebuiltTRG", false);
  p.preGraphPath  = s("trg.prebuiltTRGPath");
  p.isUpdate                 = b("trg.isUpdate", false);
  p.expandDist               = f("trg.expandDist", 0.6f);
  p.robotSize                = f("trg.robotSize", 0.3f);
  p.sampleNum                = has("trg.sampleNum") ? std::stoi(y.at("trg.sampleNum")) : 20;
  p.heightThreshold          = f("trg.heightThreshold", 0.15f);
  p.collisionThreshold       = f("trg.collisionThreshold", 0.2f);
  p.updateCollisionThreshold = f("trg.updateCollisionThreshold", 0.2f);
  p.safetyFactor             = f("trg.safetyFactor", 1.0f);
  p.goal_tolerance           = f("trg.goalTolerance", 0.8f);
  return p;
}

std::vector<float> load_pcd_xyz(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("trg_b200: Failed to load prebuilt map: " + path);
  std::vector<std::string> fields;
  std::vector<int> sizes, counts;
  std::vector<char> types;
  int64_t points = -1, width = 0, height = 1;
  std::string data_kind, line;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line[0] == '#') continue;
    std::istringstream is(line);
    std::string tag;
    is >> tag;
    std::string tok;
    if (tag == "FIELDS") while (is >> tok) fields.push_back(tok);
    else if (tag == "SIZE") while (is >> tok) sizes.push_back(std::stoi(tok));
    else if (tag == "TYPE") while (is >> tok) types.push_back(tok[0]);
    else if (tag == "COUNT") while (is >> tok) counts.push_back(std::stoi(tok));
    else if (tag == "WIDTH") is >> width;
    else if (tag == "HEIGHT") is >> height;
    else if (tag == "POINTS") is >> points;
    else if (tag == "DATA") { is >> data_kind; break; }
  }
  if (points < 0) points = width * height;
  if (counts.empty()) counts.assign(fields.size(), 1);
  if (fields.empty() || sizes.size() != fields.size() || types.size() != fields.size())
    throw std::runtime_error("trg_b200: malformed PCD header: " + path);
  int off[3] = {-1, -1, -1}, col[3] = {-1, -1, -1};
  int stride = 0, ncol = 0;
  std::vector<int> f_off(fields.size());
  for (size_t i = 0; i < fields.size(); ++i) {
    f_off[i] = stride;
    for (int k = 0; k < 3; ++k)
      if (fields[i] == std::string(1, "xyz"[k])) {
        if (sizes[i] != 4 || types[i] != 'F') throw std::runtime_error("trg_b200: PCD x/y/z must be 4-byte floats");
        off[k] = stride;
        col[k] = ncol;
      }
    stride += sizes[i] * counts[i];
    ncol += counts[i];
  }
  if (off[0] < 0 || off[1] < 0 || off[2] < 0) throw std::runtime_error("trg_b200: PCD has no x y z fields");
  std::vector<float> xyz((size_t)points * 3);
  if (data_kind == "ascii") {
    std::vector<double> row(ncol);
    for (int64_t i = 0; i < points; ++i) {
      for (int c = 0; c < ncol; ++c)
        if (!(f >> row[c])) throw std::runtime_error("trg_b200: truncated ascii PCD");
      for (int k = 0; k < 3; ++k) xyz[3 * i + k] = (float)row[col[k]];
    }
  } else if (data_kind == "binary") {
    std::vector<char> buf((size_t)points * stride);
    f.read(buf.data(), (std::streamsize)buf.size());
    if ((size_t)f.gcount() != buf.size()) throw std::runtime_error("trg_b200: truncated binary PCD");
    for (int64_t i = 0; i < points; ++i)
      for (int k = 0; k < 3; ++k) std::memcpy(&xyz[3 * i + k], buf.data() + i * stride + off[k], 4);
  } else if (data_kind == "binary_compressed") {
    uint32_t csize = 0, usize = 0;
    f.read(reinterpret_cast<char*>(&csize), 4);
    f.read(reinterpret_cast<char*>(&usize), 4);
    std::vector<uint8_t> comp(csize), raw(usize);
    f.read(reinterpret_cast<char*>(comp.data()), csize);
    if ((uint32_t)f.gcount() != csize || lzf_decompress(comp.data(), csize, raw.data(), usize) != usize)
      throw std::runtime_error("trg_b200: corrupt binary_compressed PCD");
    // the payload is stored field by field (structure of arrays)
    for (int k = 0; k < 3; ++k) {
      size_t field_i = 0;
      for (size_t i = 0; i < fields.size(); ++i)
        if (f_off[i] == off[k]) field_i = i;
      const size_t base = (size_t)f_off[field_i] * points;
      for (int64_t i = 0; i < points; ++i) std::memcpy(&xyz[3 * i + k], raw.data() + base + (size_t)i * 4, 4);
    }
  } else {
    throw std::runtime_error("trg_b200: unsupported PCD DATA kind '" + data_kind + "'");
  }
  return xyz;
}

void save_pcd_xyz(const std::string& path, const float* xyz, int64_t n, bool binary) {
  std::ofstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("trg_b200: cannot write " + path);
  f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
    << "WIDTH " << n << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA " << (binary ? "binary" : "ascii") << "\n";
  if (binary) {
    f.write(reinterpret_cast<const char*>(xyz), (std::streamsize)(n * 3 * sizeof(float)));
  } else {
    f.precision(9);
    for (int64_t i = 0; i < n; ++i) f << xyz[3 * i] << " " << xyz[3 * i + 1] << " " << xyz[3 * i + 2] << "\n";
  }
}

}  // namespace trg_b200
