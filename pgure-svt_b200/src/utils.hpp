// CLI-side helpers with the reference's behaviour: Print/PrintFixed, ElapsedSeconds, StrToBool and the `.svt`
// parameter-file parser (reference: src/utils.hpp:27-94).  The parser's quirks are part of the file format
// (SURVEY Q25): first token = key, second token = separator, remaining tokens are concatenated WITHOUT spaces
// until a token contains '#'; a value that is the very last token of a line with nothing after it gets a
// trailing space appended.
#ifndef PGURESVT_B200_UTILS_HPP
#define PGURESVT_B200_UTILS_HPP
#include <algorithm>
#include <cctype>
#include <chrono>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <string>

namespace pguresvt
{
template <typename... Args>
void Print(std::ostream &out, Args &&...args)
{
    (out << ... << args) << std::endl;
}

template <typename... Args>
void PrintFixed(const uint32_t precision, Args &&...args)
{
    Print(std::cout, std::fixed, std::setprecision(precision), args...);
}

inline double ElapsedSeconds(std::chrono::high_resolution_clock::time_point t0, std::chrono::high_resolution_clock::time_point t1)
{
    return static_cast<double>(std::chrono::duration_cast<std::chrono::microseconds>(t1 - t0).count() * 1E-6);
}

inline bool StrToBool(std::string &s)
{
    for (auto &ch : s)
        ch = (char)std::tolower((unsigned char)ch);
    return s == "1" || s == "true";
}

inline void ParseParameters(std::istream &cfgfile, std::map<std::string, std::string> &options)
{
    std::string line;
    while (std::getline(cfgfile, line))
    {
        std::istringstream ls(line);
        std::string key, sep, value, tok;
        if (!(ls >> key) || key[0] == '#')
            continue; // blank or comment line
        const bool has_sep = static_cast<bool>(ls >> sep);
        if (!has_sep || sep == ":" || ls.get() != EOF)
        {
            while (ls >> tok)
            {
                if (tok.find('#') != std::string::npos)
                    break; // inline comment (a token glued to '#' is dropped whole)
                value += tok;
                if (!(ls >> std::ws))
                    value += " "; // token ended exactly at end of line
            }
        }
        options[key] = value;
    }
}
} // namespace pguresvt
#endif
