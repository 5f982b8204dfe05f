#include "config_map.h"

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace rg {

namespace {

std::string rstrip(std::string s) {
  while (!s.empty() && std::isspace(static_cast<unsigned char>(s.back()))) s.pop_back();
  return s;
}
std::string lstrip(const std::string& s) {
  size_t i = 0;
  while (i < s.size() && std::isspace(static_cast<unsigned char>(s[i]))) ++i;
  return s.substr(i);
}
// position of the first `c`, or of a ';' that follows whitespace, or npos
size_t findCharOrComment(const std::string& s, size_t from, char c) {
  bool wasSpace = false;
  for (size_t i = from; i < s.size(); ++i) {
    if (c != '\0' && s[i] == c) return i;
    if (wasSpace && s[i] == ';') return i;
    wasSpace = std::isspace(static_cast<unsigned char>(s[i])) != 0;
  }
  return std::string::npos;
}

}  // namespace

std::string ConfigMap::makeKey(const std::string& section, const std::string& name) {
  std::string key = section + "." + name;
  for (char& ch : key) ch = static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
  return key;
}

ConfigMap ConfigMap::fromText(const std::string& text) {
  ConfigMap cfg;
  cfg.text_ = text;
  std::string section, prevName;
  size_t pos = 0;
  const size_t kMaxPiece = 199;  // fgets(line, 200)
  while (pos < text.size()) {
    size_t eol = text.find('\n', pos);
    size_t len = (eol == std::string::npos ? text.size() : eol) - pos;
    bool whole = len <= kMaxPiece;
    std::string raw = text.substr(pos, whole ? len : kMaxPiece);
    pos += whole ? len + (eol == std::string::npos ? 0 : 1) : kMaxPiece;

    std::string trimmedRight = rstrip(raw);
    std::string line = lstrip(trimmedRight);
    bool hadLeadingSpace = line.size() != trimmedRight.size();
    if (line.empty()) continue;

    if (!prevName.empty() && hadLeadingSpace) {  // continuation of the previous value
      cfg.values_[makeKey(section, prevName)] = line;
    } else if (line[0] == ';' || line[0] == '#') {
      // comment
    } else if (line[0] == '[') {
      size_t end = findCharOrComment(line, 1, ']');
      if (end != std::string::npos && line[end] == ']') {
        section = line.substr(1, end - 1).substr(0, 49);
        prevName.clear();
      }
    } else {
      size_t eq = findCharOrComment(line, 0, '=');
      if (eq != std::string::npos && line[eq] == '=') {
        std::string name = rstrip(line.substr(0, eq));
        std::string value = lstrip(line.substr(eq + 1));
        size_t cm = findCharOrComment(value, 0, '\0');
        if (cm != std::string::npos) value = value.substr(0, cm);
        value = rstrip(value);
        prevName = name.substr(0, 49);
        cfg.values_[makeKey(section, name)] = value;
      }
    }
  }
  return cfg;
}

ConfigMap ConfigMap::fromFile(const std::string& path, bool* ok) {
  std::ifstream in(path.c_str(), std::ios::in | std::ios::binary);
  if (ok) *ok = in.good();
  if (!in.good()) return ConfigMap();
  std::stringstream ss;
  ss << in.rdbuf();
  return fromText(ss.str());
}

bool ConfigMap::has(const std::string& section, const std::string& name) const {
  return values_.count(makeKey(section, name)) != 0;
}

std::string ConfigMap::getString(const std::string& section, const std::string& name,
                                 const std::string& dflt) const {
  auto it = values_.find(makeKey(section, name));
  return it == values_.end() ? dflt : it->second;
}

void ConfigMap::setString(const std::string& section, const std::string& name, const std::string& value) {
  values_[makeKey(section, name)] = value;
}

long ConfigMap::getInteger(const std::string& section, const std::string& name, long dflt) const {
  std::string v = getString(section, name, "");
  char* end = nullptr;
  long n = std::strtol(v.c_str(), &end, 0);
  return end > v.c_str() ? n : dflt;
}

float ConfigMap::getFloat(const std::string& section, const std::string& name, float dflt) const {
  std::string v = getString(section, name, "");
  char* end = nullptr;
  float f = std::strtof(v.c_str(), &end);  // float on purpose, see header
  return end > v.c_str() ? f : dflt;
}

bool ConfigMap::getBool(const std::string& section, const std::string& name, bool dflt) const {
  std::string v = getString(section, name, "");
  if (v.empty()) return dflt;
  bool val = dflt;
  if (v == "1" || v == "yes" || v == "true" || v == "on") val = true;
  if (v == "0" || v == "no" || v == "false" || v == "off") val = false;
  return val;
}

}  // namespace rg
