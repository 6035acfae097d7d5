// settings.hpp -- INI reader with the reference's Settings / Section / Entry interface (include/settings.h:47-100,
// src/settings.cpp:61-140): `[SECTION]`, `key = value`, `;` / `#` comment lines, and continuation lines (a line
// without '=' is appended, newline-separated, to the previous entry's value).  Header-only.
#pragma once
#include <istream>
#include <map>
#include <string>

namespace RGBID_SLAM {

inline std::string trim(const std::string& src, const char* delims = " \t\r\n")
{
  std::string::size_type last = src.find_last_not_of(delims);
  if (last == std::string::npos) return std::string();
  std::string::size_type first = src.find_first_not_of(delims);
  return src.substr(first, last - first + 1);
}

class Entry {
 public:
  Entry(std::string name = "", std::string value = "") : name_(std::move(name)), value_(std::move(value)) {}
  std::string getName() const { return name_; }
  std::string getValue() const { return value_; }
  void setValue(std::string v) { value_ = std::move(v); }

 private:
  std::string name_, value_;
};

class Section {
 public:
  Section(std::string name = "") : name_(std::move(name)) {}
  void addEntry(Entry& e) { entries_[e.getName()] = e; }
  bool getEntry(const std::string& name, Entry& out) const
  {
    auto it = entries_.find(name);
    if (it == entries_.end()) return false;
    out = it->second;
    return true;
  }
  std::string getName() const { return name_; }
  std::map<std::string, Entry> entries_;

 private:
  std::string name_;
};

class Settings {
 public:
  Settings() {}
  explicit Settings(std::istream& in) { load(in); }
  void load(std::istream& in)
  {
    std::string line, section, entry_name, entry_value;
    while (std::getline(in, line)) {
      line = trim(line);
      if (line.empty() || line[0] == '#' || line[0] == ';') continue;
      if (line[0] == '[') {
        section = trim(line.substr(1, line.find(']') - 1));
        Section s(section);
        addSection(s);
        continue;
      }
      std::string::size_type eq = line.find('=');
      if (eq != std::string::npos) {
        entry_name = trim(line.substr(0, eq));
        entry_value = trim(line.substr(eq + 1));
        if (!entry_name.empty()) {
          Entry e(entry_name, entry_value);
          sections_[section].addEntry(e);
        }
      } else if (!entry_name.empty()) {
        entry_value += '\n';
        entry_value += line;
        sections_[section].entries_[entry_name].setValue(entry_value);
      }
    }
  }
  void addSection(Section& s) { sections_[s.getName()] = s; }
  bool getSection(const std::string& name, Section& out) const
  {
    auto it = sections_.find(name);
    if (it == sections_.end()) return false;
    out = it->second;
    return true;
  }

 private:
  std::map<std::string, Section> sections_;
};

}  // namespace RGBID_SLAM
