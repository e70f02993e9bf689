// cli_b200.hpp -- the few lines of option parsing the three server drivers share.  The reference uses
// the vendored cxxopts (include/cxxopts.h, third party); only its observable behaviour for the
// flags the servers define is kept: "--name value", "-x value", boolean switches, `--help`, and
// "error parsing options: ..." + exit(1) for an unknown option or a missing argument
// (server_jpeg.cpp:38-71).
#pragma once
#include <cstdlib>
#include <iostream>
#include <map>
#include <set>
#include <string>

namespace cli_b200 {

struct Spec {
    std::map<std::string, std::string> alias;  // "-f" -> "file"
    std::set<std::string> valued, switches;    // long names
};

struct Parsed {
    std::map<std::string, std::string> value;
    std::set<std::string> on;
    bool has(const std::string &k) const { return value.count(k) || on.count(k); }
    std::string str(const std::string &k, const std::string &dflt) const { return value.count(k) ? value.at(k) : dflt; }
    int num(const std::string &k, int dflt) const { return value.count(k) ? std::atoi(value.at(k).c_str()) : dflt; }
    double real(const std::string &k, double dflt) const { return value.count(k) ? std::atof(value.at(k).c_str()) : dflt; }
};

inline Parsed parse(int argc, const char **argv, const Spec &spec)
{
    Parsed p;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i], name;
        if (a.rfind("--", 0) == 0) name = a.substr(2);
        else if (spec.alias.count(a)) name = spec.alias.at(a);
        else name = "";
        if (spec.switches.count(name)) {
            p.on.insert(name);
        } else if (spec.valued.count(name)) {
            if (i + 1 >= argc) {
                std::cout << "error parsing options: Option '" << name << "' is missing an argument" << std::endl;
                std::exit(1);
            }
            p.value[name] = argv[++i];
        } else {
            std::cout << "error parsing options: Option '" << a << "' does not exist" << std::endl;
            std::exit(1);
        }
    }
    return p;
}

}  // namespace cli_b200
