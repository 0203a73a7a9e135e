// options.hpp -- command-line parsing for the knn_rms / knn_data tools.
//
// The reference declares its options with boost::program_options (knn_rms.cpp:72-91,
// knn_data.cpp:68-87); Boost is not a dependency here, so this is a small parser that accepts the
// same spellings: "--name value", "--name=value", "-n value", "-nvalue", unambiguous long-option
// prefixes, po::value<bool> options that TAKE a value (true/false/1/0/yes/no/on/off) and
// po::bool_switch options that do not.
#pragma once
#include <cstdlib>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace mdsctk_cli {

class Options {
public:
    enum Kind { VALUE, SWITCH };
    struct Opt {
        std::string long_name;
        char short_name;
        Kind kind;
        std::string help, value, default_text;
        bool seen = false, has_default = false;
    };

    void add(const std::string &long_name, char short_name, Kind kind, const std::string &help,
             const std::string &default_value = "", bool has_default = false)
    {
        Opt o;
        o.long_name = long_name; o.short_name = short_name; o.kind = kind; o.help = help;
        o.value = default_value; o.default_text = default_value; o.has_default = has_default;
        opts_.push_back(o);
    }

    // Throws std::runtime_error on unknown / ambiguous options or missing arguments.
    void parse(int argc, char **argv)
    {
        for (int i = 1; i < argc; ++i) {
            std::string a = argv[i];
            Opt *o = nullptr;
            std::string inline_val;
            bool has_inline = false;
            if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
                std::string name = a.substr(2);
                size_t eq = name.find('=');
                if (eq != std::string::npos) { inline_val = name.substr(eq + 1); name = name.substr(0, eq); has_inline = true; }
                o = find_long(name);
            } else if (a.size() >= 2 && a[0] == '-' && a[1] != '-') {
                o = find_short(a[1]);
                if (a.size() > 2) { inline_val = a.substr(2); has_inline = true; if (!inline_val.empty() && inline_val[0] == '=') inline_val.erase(0, 1); }
            } else {
                throw std::runtime_error("too many positional options have been specified on the command line");
            }
            o->seen = true;
            if (o->kind == SWITCH) {
                if (has_inline) throw std::runtime_error("option '--" + o->long_name + "' does not take any arguments");
                o->value = "1";
            } else {
                if (has_inline) o->value = inline_val;
                else if (i + 1 < argc) o->value = argv[++i];
                else throw std::runtime_error("the required argument for option '--" + o->long_name + "' is missing");
            }
        }
    }

    bool count(const std::string &name) const { return get(name).seen; }
    std::string str(const std::string &name) const { return get(name).value; }
    int integer(const std::string &name) const
    {
        const std::string v = get(name).value;
        char *end = nullptr;
        long r = std::strtol(v.c_str(), &end, 10);
        if (v.empty() || *end) throw std::runtime_error("the argument ('" + v + "') for option '--" + name + "' is invalid");
        return (int)r;
    }
    bool boolean(const std::string &name) const
    {
        std::string v = get(name).value;
        for (auto &c : v) c = (char)std::tolower((unsigned char)c);
        if (v == "1" || v == "true" || v == "yes" || v == "on") return true;
        if (v == "0" || v == "false" || v == "no" || v == "off" || v.empty()) return false;
        throw std::runtime_error("the argument ('" + v + "') for option '--" + name + "' is invalid");
    }

    void print(std::ostream &os, const std::string &title) const
    {
        os << title << ":" << std::endl;
        for (const auto &o : opts_) {
            std::ostringstream l;
            l << "  -" << o.short_name << " [ --" << o.long_name << " ]";
            if (o.kind == VALUE) { l << " arg"; if (o.has_default) l << " (=" << o.default_text << ")"; }
            std::string s = l.str();
            if (s.size() < 38) s.append(38 - s.size(), ' '); else s += "\n" + std::string(38, ' ');
            os << s << o.help << std::endl;
        }
    }

private:
    std::vector<Opt> opts_;
    const Opt &get(const std::string &name) const
    {
        for (const auto &o : opts_) if (o.long_name == name) return o;
        throw std::logic_error("undeclared option " + name);
    }
    Opt *find_long(const std::string &name)
    {
        Opt *hit = nullptr;
        for (auto &o : opts_) {
            if (o.long_name == name) return &o;
            if (o.long_name.compare(0, name.size(), name) == 0) {
                if (hit) throw std::runtime_error("option '--" + name + "' is ambiguous");
                hit = &o;
            }
        }
        if (!hit) throw std::runtime_error("unrecognised option '--" + name + "'");
        return hit;
    }
    Opt *find_short(char c)
    {
        for (auto &o : opts_) if (o.short_name == c) return &o;
        throw std::runtime_error(std::string("unrecognised option '-") + c + "'");
    }
};

// mdsctk.cpp:269-283 copyright(): the banner every tool prints first.
inline void banner(const char *program_name)
{
    std::cout << std::endl;
    std::cout << "   MDSCTK 1.2 - " << program_name << " (B200 build)" << std::endl;
    std::cout << "   Copyright (C) 2013 Joshua L. Phillips" << std::endl;
    std::cout << "   MDSCTK comes with ABSOLUTELY NO WARRANTY; see LICENSE for details." << std::endl;
    std::cout << "   This is free software, and you are welcome to redistribute it" << std::endl;
    std::cout << "   under certain conditions; see README.md for details." << std::endl;
    std::cout << std::endl;
}

}  // namespace mdsctk_cli
