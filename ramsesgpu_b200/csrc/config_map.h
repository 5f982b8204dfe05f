// ConfigMap: .ini reader with the reference's lookup semantics.
//
// Mirrors the behaviour (not the code) of the reference's INIReader + ConfigMap
// (src/utils/config/inih/ini.cpp:71-140, inih/INIReader.cpp:33-101, ConfigMap.cpp:41-87):
//   * keys are "section.name", lower-cased; later duplicates overwrite;
//   * '#' / ';' full-line comments, " ;" inline comments, continuation lines;
//   * lines are consumed in 199-character pieces (the reference reads with fgets(line, 200));
//   * getFloat parses with strtof and returns FLOAT even when the solver runs in double
//     (gamma0=1.66 is (double)1.66f in the reference; parity depends on it);
//   * getBool accepts 1/yes/true/on and 0/no/false/off.
#pragma once
#include <map>
#include <string>

namespace rg {

class ConfigMap {
 public:
  ConfigMap() = default;
  static ConfigMap fromText(const std::string& text);
  static ConfigMap fromFile(const std::string& path, bool* ok = nullptr);

  std::string getString(const std::string& section, const std::string& name, const std::string& dflt) const;
  long getInteger(const std::string& section, const std::string& name, long dflt) const;
  float getFloat(const std::string& section, const std::string& name, float dflt) const;
  bool getBool(const std::string& section, const std::string& name, bool dflt) const;
  void setString(const std::string& section, const std::string& name, const std::string& value);
  bool has(const std::string& section, const std::string& name) const;
  const std::map<std::string, std::string>& values() const { return values_; }
  const std::string& text() const { return text_; }

 private:
  static std::string makeKey(const std::string& section, const std::string& name);
  std::map<std::string, std::string> values_;
  std::string text_;
};

}  // namespace rg
