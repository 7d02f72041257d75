// PGURE-SVT command-line tool — drop-in for the reference's CLI (src/PGURE-SVT.cpp:18-246): same usage
// (`PGURE-SVT paramfile`), same `.svt` keys and defaults, same input/output naming (<stem>.tif →
// <stem>-CLEANED.tif, 16-bit), same four timing lines.  The denoising itself is PGURESVT<uint16_t,double>()
// → C ABI → CUDA.  Divergences, all documented in DESIGN.md: TIFF I/O is an in-tree baseline reader/writer
// (libtiff is absent), 8-bit pages are widened properly, `optimize_pgure : true` passes "no start value"
// (-1) instead of the reference's 0.0, which makes NLopt throw.
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "pguresvt.hpp"
#include "tiff_min.hpp"
#include "utils.hpp"

int main(int argc, char **argv)
{
    std::chrono::high_resolution_clock::time_point t0, t1;
    pguresvt::Print(std::cout, "PGURE-SVT Denoising (B200)\n", "Reference: T. Furnival et al., Ultramicroscopy 178 (2017)\n");
    if (argc != 2)
    {
        pguresvt::Print(std::cout, "  Usage: ./PGURE-SVT paramfile");
        return -1;
    }
    std::map<std::string, std::string> opts;
    std::ifstream paramFile(argv[1], std::ios::in);
    pguresvt::ParseParameters(paramFile, opts);
    if (opts.count("filename") == 0 || opts.count("start_frame") == 0 || opts.count("end_frame") == 0)
    {
        pguresvt::Print(std::cerr, "**ERROR**\n", "Required parameters not specified\n",
                        "You must specify 'filename', 'start_frame' and 'end_frame'\n");
        return -1;
    }
    auto has = [&](const char *k) { return opts.count(k) == 1; };
    auto geti = [&](const char *k, int d) { return has(k) ? std::stoi(opts.at(k)) : d; };
    auto getd = [&](const char *k, double d) { return has(k) ? std::stod(opts.at(k)) : d; };
    auto getb = [&](const char *k, bool d) { return has(k) ? pguresvt::StrToBool(opts.at(k)) : d; };

    const std::string filename = opts.at("filename");
    const std::string filestem = filename.substr(0, filename.find_last_of("."));
    const uint32_t startImage = (uint32_t)std::stoi(opts.at("start_frame"));
    const uint32_t endImage = (uint32_t)std::stoi(opts.at("end_frame"));
    const uint32_t nImages = endImage - startImage + 1;

    const uint32_t blockSize = geti("patch_size", 4);
    const uint32_t trajLength = geti("trajectory_length", 15);
    const uint32_t motionWindow = geti("motion_neighbourhood", 7);
    const uint32_t medianSize = geti("median_filter", 5);
    const uint32_t blockOverlap = geti("patch_overlap", 1);
    const uint32_t noiseMethod = geti("noise_method", 4);
    const uint32_t maxIter = geti("max_iter", 1000);
    const int nJobs = geti("n_jobs", -1);
    const double alpha = getd("noise_alpha", -1.);
    const double mu = getd("noise_mu", -1.);
    const double sigma = getd("noise_sigma", -1.);
    const double hotPixelThreshold = has("hot_pixel") ? std::stoi(opts.at("hot_pixel")) : -1.0;
    const int randomSeed = geti("random_seed", -1);
    const bool expWeighting = getb("exponential_weighting", true);
    const bool normalizeImg = getb("normalize", false);
    const bool motionEstimation = getb("motion_estimation", true);
    const bool optPGURE = getb("optimize_pgure", true);
    double lambda = -1.0; // reference passes 0.0 here, which NLopt rejects: treated as "unset"
    if (!optPGURE)
    {
        if (has("lambda"))
            lambda = std::stod(opts.at("lambda"));
        else
        {
            pguresvt::Print(std::cerr, "**ERROR**\nPGURE optimization is turned OFF but ", "no lambda specified in parameter file\n");
            return -1;
        }
    }
    const double tol = 1E-7; // the reference parses `tolerance` into a shadowed local, i.e. ignores it (PGURE-SVT.cpp:93-99)

    t0 = std::chrono::high_resolution_clock::now();
    const std::string inFilename = filestem + ".tif";
    if (!std::ifstream(inFilename.c_str()))
    {
        pguresvt::Print(std::cerr, "**ERROR**\nFile ", inFilename, " not found\n");
        return -1;
    }
    tiffmin::Reader tif;
    if (!tif.open(inFilename))
    {
        pguresvt::Print(std::cerr, "**ERROR**\nCould not read ", inFilename, ": ", tif.error, "\n");
        return -1;
    }
    const uint32_t W = tif.page(0).width, H = tif.page(0).height;
    for (uint32_t d = 0; d < tif.n_pages(); d++) // every page must have the geometry of the first (the cube is sized from page 0)
        if (tif.page(d).width != W || tif.page(d).height != H)
        {
            pguresvt::Print(std::cerr, "**ERROR**\nTIFF page ", d + 1, " is ", tif.page(d).width, "x", tif.page(d).height, ", page 1 is ", W, "x", H, "\n");
            return -1;
        }
    const uint16_t depth = tif.page(0).bits;
    if (W != H)
    {
        pguresvt::Print(std::cerr, "**ERROR**\nFrame dimensions are not square, got ", W, "x", H, "\n");
        return -1;
    }
    if (depth != 8 && depth != 16)
    {
        pguresvt::Print(std::cerr, "**ERROR**\nImages must be 8-bit or 16-bit, got ", depth, "-bit depth \n");
        return -1;
    }
    // pages [start-1, end) → cube(row = y, col = x, frame): column-major memory = transposed scanline buffer
    std::vector<uint32_t> take;
    for (uint32_t d = 0; d < tif.n_pages(); d++)
        if (d >= startImage - 1 && d < endImage)
            take.push_back(d);
    if (nImages > take.size())
    {
        pguresvt::Print(std::cerr, "**ERROR**\n Sequence only has ", take.size(), " frames, expected ", nImages, "\n");
        return -1;
    }
    arma::Cube<uint16_t> inputSeq(H, W, take.size());
    {
        std::vector<uint16_t> buf((size_t)W * H);
        for (size_t k = 0; k < take.size(); k++)
        {
            if (!tif.read_page(take[k], buf.data()))
            {
                pguresvt::Print(std::cerr, "**ERROR**\nCould not read page ", take[k], ": ", tif.error, "\n");
                return -1;
            }
            uint16_t *dst = inputSeq.slice_memptr(k);
            for (uint32_t y = 0; y < H; y++)
                for (uint32_t x = 0; x < W; x++)
                    dst[y + (size_t)H * x] = buf[(size_t)y * W + x];
        }
    }
    t1 = std::chrono::high_resolution_clock::now();
    pguresvt::PrintFixed(4, "TIFF import:    ", std::setw(10), pguresvt::ElapsedSeconds(t0, t1), " seconds");

    if (hotPixelThreshold >= 0.0)
    {
        t0 = std::chrono::high_resolution_clock::now();
        const int rc = pguresvt_hotpixel_u16(inputSeq.memptr(), H, W, (uint32_t)inputSeq.n_slices, hotPixelThreshold,
                                             std::getenv("PGURESVT_DEVICE") ? std::atoi(std::getenv("PGURESVT_DEVICE")) : 0);
        if (rc != 0)
        {
            pguresvt::Print(std::cerr, "**ERROR**\nOutlier filter failed: ", pguresvt_last_error(), "\n");
            return -1;
        }
        t1 = std::chrono::high_resolution_clock::now();
        pguresvt::PrintFixed(4, "Outlier filter: ", std::setw(10), pguresvt::ElapsedSeconds(t0, t1), " seconds");
    }

    t0 = std::chrono::high_resolution_clock::now();
    arma::cube cleanSeq;
    arma::mat res;
    const uint32_t result = PGURESVT(cleanSeq, res, inputSeq, trajLength, blockSize, blockOverlap, motionWindow, (int64_t)medianSize,
                                     noiseMethod, maxIter, (int64_t)nJobs, (int64_t)randomSeed, optPGURE, expWeighting, motionEstimation,
                                     lambda, alpha, mu, sigma, tol);
    t1 = std::chrono::high_resolution_clock::now();
    if (result != 0)
    {
        pguresvt::Print(std::cerr, "**ERROR**\nPGURE-SVT failed (code ", result, "): ", pguresvt_last_error(), "\n");
        return (int)result;
    }
    pguresvt::PrintFixed(4, "PGURE-SVT:      ", std::setw(10), pguresvt::ElapsedSeconds(t0, t1), " seconds");

    t0 = std::chrono::high_resolution_clock::now();
    if (normalizeImg)
    { // 65535 * (x - min) / (max - min)   (PGURE-SVT.cpp:200-203)
        const double lo = cleanSeq.min(), hi = cleanSeq.max();
        double *d = cleanSeq.memptr();
        for (arma::uword k = 0; k < cleanSeq.n_elem; k++)
            d[k] = 65535 * (d[k] - lo) / (hi - lo);
    }
    const std::string outFilename = filestem + "-CLEANED.tif";
    tiffmin::Writer out;
    if (!out.open(outFilename))
    {
        pguresvt::Print(std::cerr, "**ERROR**\nFile ", outFilename, " could not be written\n");
        return -1;
    }
    {
        std::vector<uint16_t> buf((size_t)W * H);
        for (uint32_t t = 0; t < nImages; t++)
        {
            const double *src = cleanSeq.slice_memptr(t);
            for (uint32_t y = 0; y < H; y++)
                for (uint32_t x = 0; x < W; x++)
                {
                    // arma::conv_to<Cube<uint16_t>>: truncation, negatives clamp to 0
                    const double v = src[y + (size_t)H * x];
                    buf[(size_t)y * W + x] = (v < 0.0) ? (uint16_t)0 : (uint16_t)(long long)v;
                }
            out.write_page(buf.data(), W, H, (uint16_t)t, (uint16_t)nImages);
        }
        out.close();
    }
    t1 = std::chrono::high_resolution_clock::now();
    pguresvt::PrintFixed(4, "TIFF export:    ", std::setw(10), pguresvt::ElapsedSeconds(t0, t1), " seconds\n");
    pguresvt::Print(std::cerr, "Output file:    ", outFilename, "\n");
    return (int)result;
}
