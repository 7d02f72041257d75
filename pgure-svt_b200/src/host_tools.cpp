// Small host-only tool used by the CPU tests to exercise the CLI's support code without a GPU:
//   host_tools parse <file.svt>            → one line per option:  key=[value]
//   host_tools tiffinfo <file.tif>         → pages width height bits + per-page 64-bit sums
//   host_tools tiffcopy <in.tif> <out.tif> → re-writes all pages as 16-bit (reader → writer round trip)
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <vector>

#include "tiff_min.hpp"
#include "utils.hpp"

int main(int argc, char **argv)
{
    if (argc < 3)
        return 2;
    const std::string cmd = argv[1];
    if (cmd == "parse")
    {
        std::ifstream f(argv[2]);
        std::map<std::string, std::string> o;
        pguresvt::ParseParameters(f, o);
        for (auto &kv : o)
            std::cout << kv.first << "=[" << kv.second << "]\n";
        return 0;
    }
    tiffmin::Reader r;
    if (!r.open(argv[2]))
    {
        std::cerr << r.error << "\n";
        return 1;
    }
    if (cmd == "tiffinfo")
    {
        std::cout << r.n_pages() << " " << r.page(0).width << " " << r.page(0).height << " " << r.page(0).bits << "\n";
        std::vector<uint16_t> buf((size_t)r.page(0).width * r.page(0).height);
        for (size_t p = 0; p < r.n_pages(); p++)
        {
            if (!r.read_page(p, buf.data()))
                return 1;
            unsigned long long s = 0;
            for (uint16_t v : buf)
                s += v;
            std::cout << s << "\n";
        }
        return 0;
    }
    if (cmd == "tiffcopy" && argc >= 4)
    {
        tiffmin::Writer w;
        if (!w.open(argv[3]))
            return 1;
        std::vector<uint16_t> buf((size_t)r.page(0).width * r.page(0).height);
        for (size_t p = 0; p < r.n_pages(); p++)
        {
            if (!r.read_page(p, buf.data()))
                return 1;
            w.write_page(buf.data(), r.page(0).width, r.page(0).height, (uint16_t)p, (uint16_t)r.n_pages());
        }
        w.close();
        return 0;
    }
    return 2;
}
