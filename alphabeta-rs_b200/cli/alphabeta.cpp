// alphabeta — command-line front end with the flags, console block and output files of the reference's
// `alphabeta` binary (src/cli/alphabeta.rs:8-38, src/arguments.rs:96-114, src/alphabeta.rs:23-59), on top of libabfit.
//
//   alphabeta -n nodelist.txt -e edgelist.txt [-i 1000] [-p 0.99] [-o .] [--seed N] [--device 0 | --devices 0-7]
//
// Differences that are deliberate: the random starts / resamples come from a seed (default 0xAB0B200,
// `--seed` is an extra flag; the reference uses an unseeded thread_rng), the two progress bars advance per stage
// of the batched call instead of per start (cli/progress.h), bootstrap.png is drawn by the library's own rasteriser
// (same content as src/plot.rs:84-137, not the same pixels), and a window on which the reference panics reports an
// error instead.
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "../../include/abfit.h"
#include "progress.h"

static bool exists(const std::string &p)
{
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

static std::string f64s(double v)
{
    char buf[512];
    abfit_format_f64(v, buf, sizeof buf);
    return buf;
}

int main(int argc, char **argv)
{
    long iterations = 1000;
    std::string edges = "./edgelist.txt", nodes = "./nodelist.txt", output = ".";
    double filter = 0.99;
    unsigned long long seed = 0xAB0B200ull;
    int device = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&](const char *name) -> const char * {
            if (i + 1 >= argc) {
                std::fprintf(stderr, "error: a value is required for '%s'\n", name);
                std::exit(2);
            }
            return argv[++i];
        };
        if (a == "-i" || a == "--iterations") iterations = std::atol(val("--iterations"));
        else if (a == "-e" || a == "--edges") edges = val("--edges");
        else if (a == "-n" || a == "--nodes") nodes = val("--nodes");
        else if (a == "-p" || a == "--posterior-max-filter") filter = std::atof(val("--posterior-max-filter"));
        else if (a == "-o" || a == "--output") output = val("--output");
        else if (a == "--seed") seed = std::strtoull(val("--seed"), nullptr, 0);
        else if (a == "--device") device = std::atoi(val("--device"));
        else if (a == "--devices") device = std::atoi(val("--devices"));  // one pedigree = one window: its starts and replicates stay on the first listed GPU
        else if (a == "-h" || a == "--help") {
            std::printf("Usage: alphabeta [OPTIONS]\n\nOptions:\n"
                        "  -i, --iterations <ITERATIONS>  Number of iterations to run for Nelder-Mead optimization [default: 1000]\n"
                        "  -e, --edges <EDGES>            Relative or absolute path to an edgelist [default: ./edgelist.txt]\n"
                        "  -n, --nodes <NODES>            Relative or absolute path to a nodelist [default: ./nodelist.txt]\n"
                        "  -p, --posterior-max-filter <P> Minimum posterior probability for a basepair read [default: 0.99]\n"
                        "  -o, --output <OUTPUT>          Output directory, must exist [default: .]\n"
                        "      --seed <SEED>              Seed of the random starts and resamples [default: 0xAB0B200]\n"
                        "      --device <N>               CUDA device [default: 0]\n");
            return 0;
        } else {
            std::fprintf(stderr, "error: unexpected argument '%s' found\n", a.c_str());
            return 2;
        }
    }
    // validators of src/arguments.rs:116-140
    for (const std::string *f : {&edges, &nodes}) {
        if (!exists(*f)) {
            std::fprintf(stderr, "error: Please provide a valid file path. By default, we fill try %s, which does not exist.\n", f->c_str());
            return 2;
        }
        std::printf("Using default file: %s\n", f->c_str());
    }
    if (!exists(output)) {
        std::fprintf(stderr, "error: Please provide a valid output directory. By default, we fill try %s, which does not exist.\n", output.c_str());
        return 2;
    }
    if (iterations <= 0) {
        std::fprintf(stderr, "error: --iterations must be positive\n");
        return 2;
    }
    abfit_ctx *ctx = nullptr;
    if (abfit_ctx_create(device, &ctx)) {
        std::printf("Error: %s\n", abfit_last_error());
        return 1;
    }
    std::printf("Building pedigree...\n");
    abfit_pedigree *ped = nullptr;
    if (abfit_pedigree_build(ctx, nodes.c_str(), edges.c_str(), filter, &ped)) {
        std::printf("Error: Error while building pedigree: %s\n", abfit_last_error());
        return 1;
    }
    std::fputs(abfit_pedigree_warnings(ped), stdout);
    int32_t n_pairs = 0;
    double p0uu = 0.0;
    abfit_pedigree_info(ped, &n_pairs, &p0uu, nullptr, nullptr);
    const double *rows = abfit_pedigree_rows(ped);
    if (n_pairs <= 0) {
        std::printf("Error: Model failed: the pedigree has no pairs\n");
        return 1;
    }
    double max_div = rows[3];
    for (int i = 1; i < n_pairs; ++i) max_div = rows[4 * i + 3] > max_div ? rows[4 * i + 3] : max_div;  // src/ab_neutral.rs:25-29
    const int n = (int)iterations;
    std::vector<double> simplices((size_t)n * 20), rows_out((size_t)n * 7), pred(n_pairs), resid(n_pairs), analysis(32);
    std::vector<int32_t> idx((size_t)n * n_pairs);
    abfit_gen_start_simplices(seed, 0, n, max_div, simplices.data());
    abfit_gen_resample_idx(seed, 0, n, n_pairs, idx.data());
    abfit_problem prob{rows, n_pairs, p0uu, p0uu, 1.0};  // eqp = p0uu, eqp_weight = 1 (src/alphabeta.rs:33-37)
    abfit_fit best;
    int32_t status = 0;
    progress::Bar pb_neutral("ABNeutral", (unsigned long long)n, false), pb_boot("BootModel", (unsigned long long)n, false);  // src/progress.rs:5-23
    pb_neutral.tick();
    const int rc = abfit_alphabeta_batch(ctx, &prob, 1, n, simplices.data(), n, idx.data(), seed, 0, 10000, 1000, DBL_EPSILON, 0,
                                         &best, pred.data(), resid.data(), &status, rows_out.data(), analysis.data());
    if (rc || status) {
        std::printf("Error: Model failed: %s\n", rc ? abfit_last_error() : "NaN in the pedigree or in every fit (the reference panics here)");
        return 1;
    }
    pb_neutral.finish();
    pb_boot.finish();
    char text[8192];
    abfit_format_analysis(analysis.data(), text, sizeof text);
    std::printf("##########\nResults:\n\n");
    std::printf("Model:\n\tAlpha: %s\n\tBeta: %s\n\tWeight: %s\n\tIntercept: %s\n", f64s(best.theta[0]).c_str(), f64s(best.theta[1]).c_str(),
                f64s(best.theta[2]).c_str(), f64s(best.theta[3]).c_str());
    std::printf("%s\n", text);
    std::printf("Estimated steady state %s\n", f64s(abfit_steady_state(best.theta[0], best.theta[1])).c_str());
    std::printf("Observed steady state methylation %s\n", f64s(1.0 - p0uu).c_str());
    std::printf("##########\n");
    const std::string pp = output + "/pedigree.txt", ap = output + "/analysis.txt", rp = output + "/raw.npy";
    std::printf("Writing pedigree to file: %s\n", pp.c_str());
    std::printf("Writing model to file: %s\n", ap.c_str());
    const int64_t shape[2] = {n, 7};
    if (abfit_write_pedigree(pp.c_str(), rows, n_pairs) || abfit_write_analysis(ap.c_str(), analysis.data()) ||
        abfit_write_npy_f64(rp.c_str(), rows_out.data(), 2, shape)) {
        std::printf("Error: %s\n", abfit_last_error());
        return 1;
    }
    {  // plot::bootstrap(results.column(0), results.column(1), output_dir)  (src/boot_model.rs:105-109)
        std::vector<double> alphas(n), betas(n);
        for (int i = 0; i < n; ++i) {
            alphas[i] = rows_out[(size_t)i * 7];
            betas[i] = rows_out[(size_t)i * 7 + 1];
        }
        if (abfit_plot_bootstrap((output + "/bootstrap.png").c_str(), alphas.data(), betas.data(), n)) {
            std::printf("Error: %s\n", abfit_last_error());
            return 1;
        }
    }
    abfit_pedigree_free(ped);
    abfit_ctx_destroy(ctx);
    return 0;
}
