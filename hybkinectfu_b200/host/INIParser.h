// INIParser.h — [section] key=value reader for config.ini (replaces src/utils/INIParser.{h,cpp}).
// Differences from the reference, all bug fixes: sections do not inherit the keys of
// earlier sections, a missing key is reported instead of dereferencing end(), '#'/';' comments and blank lines
// are skipped, values are trimmed.
#pragma once
#include <map>
#include <string>

class INIParser {
public:
    bool readINI(const std::string& path);
    bool has(const std::string& section, const std::string& key) const;
    // returns "" when absent (and records the miss, see missing())
    std::string getValue(const std::string& section, const std::string& key) const;
    double getNumber(const std::string& section, const std::string& key, double dflt) const;
    const std::string& missing() const { return _missing; }
    void clear() { _map.clear(); _missing.clear(); }
private:
    std::map<std::string, std::map<std::string, std::string>> _map;
    mutable std::string _missing;
};
