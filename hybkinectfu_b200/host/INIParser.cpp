#include "INIParser.h"
#include <cstdlib>
#include <fstream>

static std::string trim(const std::string& s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

bool INIParser::readINI(const std::string& path)
{
    std::ifstream in(path.c_str());
    if (!in.is_open()) return false;
    clear();
    std::string line, section;
    while (std::getline(in, line)) {
        line = trim(line);
        if (line.empty() || line[0] == '#' || line[0] == ';') continue;
        if (line[0] == '[') {
            size_t e = line.find(']');
            if (e == std::string::npos) return false;
            section = trim(line.substr(1, e - 1));
            _map[section];
            continue;
        }
        size_t eq = line.find('=');
        if (eq == std::string::npos) continue;
        _map[section][trim(line.substr(0, eq))] = trim(line.substr(eq + 1));
    }
    return true;
}

bool INIParser::has(const std::string& section, const std::string& key) const
{
    auto s = _map.find(section);
    return s != _map.end() && s->second.find(key) != s->second.end();
}

std::string INIParser::getValue(const std::string& section, const std::string& key) const
{
    auto s = _map.find(section);
    if (s != _map.end()) {
        auto k = s->second.find(key);
        if (k != s->second.end()) return k->second;
    }
    _missing += "[" + section + "]" + key + " ";
    return std::string();
}

double INIParser::getNumber(const std::string& section, const std::string& key, double dflt) const
{
    if (!has(section, key)) { _missing += "[" + section + "]" + key + " "; return dflt; }
    return std::atof(getValue(section, key).c_str());
}
