// main.cpp -- `sabreur` command line over the GPU library: the reference's driver (src/main.rs:22-198, src/cli.rs)
// with the same arguments, output layout, log lines and exit codes.
//
//   sabreur [options] <BARCODE> <FORWARD FILE> [<REVERSE FILE>]
//     -m, --mismatch <N>   maximum number of mismatches [0]          (cli.rs:32-34)
//     -o, --output <DIR>   output directory [sabreur_out]            (cli.rs:36-38)
//     -f, --format <FMT>   gz | xz | bz2 | zst                       (cli.rs:40-42)
//     -l, --level <1-9>    compression level [1]                     (cli.rs:44-46)
//         --force          reuse the output directory                (cli.rs:48-50)
//     -q, --quiet          decrease verbosity                        (cli.rs:52-54)
//   additions (not in the reference; defaults keep its behaviour):
//         --gpus <N>       deal batches round-robin to N GPUs [1]
//         --batch-mb <N>   size of one GPU batch in MiB [256]
//         --timing         print where the wall time went
//         --sequential-mates   paired-end: one pass after the other, like the reference (default: both at once)
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include <ftw.h>

#include "demux_host.hpp"

using namespace sabreur;

namespace {

constexpr int EX_USAGE_CLAP = 2;   // clap's exit code for argument errors
constexpr int EX_CANTCREAT = 73;   // exitcode::CANTCREAT (main.rs:55)
constexpr int EX_PANIC = 101;      // a Rust panic (short read, malformed barcode row)

FILE* g_logfile = nullptr;
bool g_quiet = false;

// fern configuration of utils.rs:15-53: file "[Y-m-d][H:M:S][target][LEVEL] msg", stdout "[H:M:S][LEVEL] msg";
// --quiet raises the filter to Warn on both
void log_msg(const char* level, int severity /*0 info, 1 warn, 2 error*/, const char* fmt, ...) {
    if (g_quiet && severity < 1) return;
    char msg[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(msg, sizeof msg, fmt, ap);
    va_end(ap);
    time_t t = time(nullptr);
    struct tm tmv;
    localtime_r(&t, &tmv);
    char d[32], h[32];
    strftime(d, sizeof d, "[%Y-%m-%d]", &tmv);
    strftime(h, sizeof h, "%H:%M:%S", &tmv);
    if (g_logfile) { fprintf(g_logfile, "%s[%s][sabreur][%s] %s\n", d, h, level, msg); fflush(g_logfile); }
    printf("[%s][%s] %s\n", h, level, msg);
    fflush(stdout);
}
#define INFO(...) log_msg("INFO", 0, __VA_ARGS__)
#define WARN(...) log_msg("WARN", 1, __VA_ARGS__)
#define ERROR(...) log_msg("ERROR", 2, __VA_ARGS__)

bool is_file(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}
bool exists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}
int rm_cb(const char* p, const struct stat*, int, struct FTW*) { return remove(p); }
bool remove_dir_all(const std::string& p) { return nftw(p.c_str(), rm_cb, 64, FTW_DEPTH | FTW_PHYS) == 0; }
bool create_dir_all(const std::string& p) {
    std::string cur;
    for (size_t i = 0; i <= p.size(); ++i) {
        if (i == p.size() || p[i] == '/') {
            if (!cur.empty() && !exists(cur) && mkdir(cur.c_str(), 0777) != 0) return false;
        }
        if (i < p.size()) cur.push_back(p[i]);
    }
    return true;
}

[[noreturn]] void usage_error(const std::string& what) {
    fprintf(stderr, "error: %s\n\nUsage: sabreur [options] <BARCODE> <FORWARD FILE> [<REVERSE FILE>]\n\nFor more information, try '--help'.\n",
            what.c_str());
    exit(EX_USAGE_CLAP);
}

void print_help() {
    puts("A barcode demultiplexer for FASTA and FASTQ files (B200-native demultiplexing path)\n\n"
         "Usage: sabreur [options] <BARCODE> <FORWARD FILE> [<REVERSE FILE>]\n\n"
         "Arguments:\n"
         "  <BARCODE>  Input barcode file\n"
         "  <FORWARD>  Input forward fastx file\n"
         "  [REVERSE]  Input reverse fastx file (optional)\n\n"
         "Options:\n"
         "  -m, --mismatch <MISMATCH>  Maximum number of mismatches [default: 0]\n"
         "  -o, --output <OUTPUT>      Output directory [default: sabreur_out]\n"
         "  -f, --format <FORMAT>      Output files compression format [gz, xz, bz2, zst]\n"
         "  -l, --level <LEVEL>        Compression level (1-9) [default: 1]\n"
         "      --force                Force reuse of output directory\n"
         "  -q, --quiet                Decrease program verbosity\n"
         "      --gpus <N>             Number of GPUs batches are dealt to [default: 1]\n"
         "      --batch-mb <N>         GPU batch size in MiB [default: 256]\n"
         "      --timing               Report where the wall time went\n"
         "      --sequential-mates     Paired-end: run the reverse pass after the forward pass (default: both at once)\n"
         "  -h, --help                 Print help\n"
         "  -V, --version              Print version\n\n"
         "Note: `sabreur -h` prints a short and concise overview while `sabreur --help` gives all details.");
}

struct Cli {
    std::string barcode, forward, reverse;
    bool has_reverse = false;
    int mismatch = 0;
    std::string output = "sabreur_out";
    bool has_format = false;
    Format format = Format::No;
    int level = 1;
    bool force = false, quiet = false, timing = false, sequential_mates = false;
    int gpus = 1;
    int batch_mb = 256;
    bool batch_mb_given = false;
};

long parse_u8(const std::string& name, const char* v) {
    char* end = nullptr;
    long x = strtol(v, &end, 10);
    if (!*v || *end || x < 0 || x > 255) usage_error("invalid value '" + std::string(v) + "' for '" + name + "'");
    return x;
}

Cli parse_cli(int argc, char** argv) {
    Cli c;
    std::vector<std::string> pos;
    auto value = [&](int& i, const std::string& flag, const char* inline_v) -> const char* {
        if (inline_v) return inline_v;
        if (i + 1 >= argc) usage_error("a value is required for '" + flag + "' but none was supplied");
        return argv[++i];
    };
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        const char* inl = nullptr;
        std::string key = a;
        if (a.rfind("--", 0) == 0) {
            size_t eq = a.find('=');
            if (eq != std::string::npos) { key = a.substr(0, eq); inl = argv[i] + eq + 1; }
        } else if (a.size() > 2 && a[0] == '-' && a[1] != '-') {  // -m2
            key = a.substr(0, 2);
            inl = argv[i] + 2;
        }
        if (key == "-h" || key == "--help") { print_help(); exit(0); }
        else if (key == "-V" || key == "--version") { puts("sabreur 0.7.0"); exit(0); }
        else if (key == "-m" || key == "--mismatch") c.mismatch = (int)parse_u8("--mismatch <MISMATCH>", value(i, key, inl));
        else if (key == "-o" || key == "--output") c.output = value(i, key, inl);
        else if (key == "-l" || key == "--level") c.level = (int)parse_u8("--level <LEVEL>", value(i, key, inl));
        else if (key == "-f" || key == "--format") {
            std::string v = value(i, key, inl);
            c.has_format = true;
            if (v == "gz") c.format = Format::Gzip;
            else if (v == "xz") c.format = Format::Lzma;
            else if (v == "bz2") c.format = Format::Bzip;
            else if (v == "zst") c.format = Format::Zstd;
            else usage_error("invalid value '" + v + "' for '--format <FORMAT>'\n  [possible values: gz, xz, bz2, zst]");
        }
        else if (key == "--force") c.force = true;
        else if (key == "-q" || key == "--quiet") c.quiet = true;
        else if (key == "--timing") c.timing = true;
        else if (key == "--sequential-mates") c.sequential_mates = true;
        else if (key == "--gpus") c.gpus = std::max(1, atoi(value(i, key, inl)));
        else if (key == "--batch-mb") { c.batch_mb = std::max(1, atoi(value(i, key, inl))); c.batch_mb_given = true; }
        else if (a.size() > 1 && a[0] == '-') usage_error("unexpected argument '" + a + "' found");
        else pos.push_back(a);
    }
    if (pos.size() < 2) usage_error("the following required arguments were not provided:\n  <BARCODE>\n  <FORWARD>");
    if (pos.size() > 3) usage_error("unexpected argument '" + pos[3] + "' found");
    c.barcode = pos[0];
    c.forward = pos[1];
    if (pos.size() == 3) { c.reverse = pos[2]; c.has_reverse = true; }
    const char* names[3] = {"<BARCODE>", "<FORWARD>", "[REVERSE]"};
    for (size_t i = 0; i < pos.size(); ++i)
        if (!is_file(pos[i])) usage_error("invalid value '" + pos[i] + "' for '" + names[i] + "': path does not exist");
    return c;
}

OutFile open_append(const std::string& dir, const std::string& name, Format ext) {  // create_relpath_from + OpenOptions (utils.rs:55-61, main.rs:86-90)
    OutFile o;
    o.path = dir + "/" + name + to_compression_ext(ext);
    o.fh = fopen(o.path.c_str(), "ab");
    if (!o.fh) throw std::runtime_error("cannot create '" + o.path + "'");
    return o;
}

}  // namespace

int main(int argc, char** argv) {
    const auto start = std::chrono::steady_clock::now();
    Cli cli = parse_cli(argc, argv);
    g_quiet = cli.quiet;
    g_logfile = fopen("sabreur.log", "a");  // fern::log_file("sabreur.log"), utils.rs:33
    try {
        const Format forward_format = which_format(cli.forward);
        Format output_format = forward_format;
        if (cli.has_format) {
            output_format = cli.format;
            INFO("Output files will %s compressed", to_compression_ext(output_format));  // (sic, main.rs:35-38)
        }
        const bool is_pe = cli.has_reverse;
        INFO("sabreur v0.7 starting up in %s mode", is_pe ? "paired-end" : "single-end");

        if (exists(cli.output)) {
            if (!cli.force) {
                ERROR("Output folder '%s' already exists. Use --force to overwrite.", cli.output.c_str());
                return EX_CANTCREAT;
            }
            INFO("Reusing directory %s", cli.output.c_str());
            if (!remove_dir_all(cli.output))
                throw std::runtime_error("Failed to remove '" + cli.output + "'. Check your permissions.");
        }
        if (!create_dir_all(cli.output)) throw std::runtime_error("Failed to create '" + cli.output + "'. Check your permissions.");

        std::ifstream bf(cli.barcode, std::ios::binary);
        std::stringstream ss;
        ss << bf.rdbuf();
        const auto fields = split_by_tab(ss.str());
        if (cli.mismatch != 0) WARN("Allowing up to %d mismatches", cli.mismatch);
        // With three or more mismatches the reference's "XXX" sentinel (demux.rs:37, 43-46) matches every read that reaches
        // it: such reads land in the unknown file while is_unk_empty stays true, and main.rs:179-181 may then delete that
        // file with reads in it.  That file-deleting quirk is not reproduced: unknown reads are kept and reported as such.
        if (cli.mismatch >= 3)
            WARN("--mismatch >= 3: the reference's 'XXX' sentinel quirk is not emulated, unmatched reads are kept in the unknown file");

        Barcode bd;
        Counts stats;
        const size_t need = is_pe ? 3 : 2;
        auto add_row = [&](const std::vector<std::string>& f, std::vector<OutFile> files) {
            // HashMap::insert: a repeated barcode keeps its place, the later row's files replace the earlier ones
            for (auto& r : bd.rows)
                if (r.barcode == f[0]) {
                    for (auto& o : r.files) fclose(o.fh);  // the file itself stays (empty), like the dropped File handle
                    r.files = std::move(files);
                    return;
                }
            bd.rows.push_back(BarcodeRow{f[0], std::move(files)});
        };
        HostOptions opt;
        opt.n_gpus = cli.gpus;
        opt.batch_bytes = (uint64_t)cli.batch_mb << 20;
        if (!cli.batch_mb_given) {
            // small inputs: do not pin and allocate 256 MiB slots for a file that decompresses to a fraction of that (the
            // allocation is most of a short run's wall time).  The estimate only has to err on the large side.
            auto estimate = [](const std::string& path, Format f) -> uint64_t {
                struct stat st;
                if (stat(path.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) return UINT64_MAX;
                const uint64_t ratio = f == Format::No ? 1 : (f == Format::Gzip || f == Format::Zstd ? 8 : 12);
                return (uint64_t)st.st_size * ratio;
            };
            uint64_t est = estimate(cli.forward, forward_format);
            if (is_pe) est = std::max(est, estimate(cli.reverse, which_format(cli.reverse)));
            if (est != UINT64_MAX) {
                const uint64_t want = std::max<uint64_t>(64ull << 20, ((est + est / 8) | ((1ull << 20) - 1)) + 1);
                opt.batch_bytes = std::min(opt.batch_bytes, want);
            }
        }
        opt.parallel_mates = !cli.sequential_mates;
        RunStats rs;
        std::string unknown_paths[2];

        if (is_pe) {
            Format reverse_format = which_format(cli.reverse);
            if (output_format != Format::No) reverse_format = output_format;  // main.rs:93-96
            for (auto& f : fields) {
                if (f.size() < need) { fprintf(stderr, "thread 'main' panicked: index out of bounds: the len is %zu but the index is %zu\n", f.size(), need - 1); return EX_PANIC; }
                add_row(f, {open_append(cli.output, f[1], output_format), open_append(cli.output, f[2], reverse_format)});
            }
            bd.unknown = {open_append(cli.output, "unknown_R1.fa", output_format), open_append(cli.output, "unknown_R2.fa", reverse_format)};
            unknown_paths[0] = bd.unknown[0].path;
            unknown_paths[1] = bd.unknown[1].path;
            auto res = pe_demux(cli.forward, cli.reverse, output_format, to_level(cli.level), bd, (uint8_t)cli.mismatch, stats, opt, &rs);
            if (!cli.quiet)
                for (auto& kv : res.first) INFO("%llu records found for barcode %s", (unsigned long long)kv.second, kv.first.c_str());
            for (auto& r : bd.rows) for (auto& o : r.files) fclose(o.fh);
            for (auto& o : bd.unknown) fclose(o.fh);
            if (res.second == "truetrue") { remove(unknown_paths[0].c_str()); remove(unknown_paths[1].c_str()); }
            else if (res.second == "truefalse") remove(unknown_paths[0].c_str());
            else if (res.second == "falsetrue") remove(unknown_paths[1].c_str());
        } else {
            for (auto& f : fields) {
                if (f.size() < need) { fprintf(stderr, "thread 'main' panicked: index out of bounds: the len is %zu but the index is %zu\n", f.size(), need - 1); return EX_PANIC; }
                add_row(f, {open_append(cli.output, f[1], output_format)});
            }
            bd.unknown = {open_append(cli.output, "unknown.fa", output_format)};
            unknown_paths[0] = bd.unknown[0].path;
            auto res = se_demux(cli.forward, output_format, to_level(cli.level), bd, (uint8_t)cli.mismatch, stats, opt, &rs);
            if (!cli.quiet)
                for (auto& kv : res.first) INFO("%llu records found for barcode %s", (unsigned long long)kv.second, kv.first.c_str());
            for (auto& r : bd.rows) for (auto& o : r.files) fclose(o.fh);
            for (auto& o : bd.unknown) fclose(o.fh);
            if (res.second) remove(unknown_paths[0].c_str());
        }

        if (!cli.quiet) {
            const auto d = std::chrono::steady_clock::now() - start;
            const long long ms = std::chrono::duration_cast<std::chrono::milliseconds>(d).count();
            const long long s = ms / 1000;
            INFO("Results saved in %s", cli.output.c_str());
            INFO("Walltime: %lldh:%lldm:%llds %lldms", s / 3600, (s / 60) % 60, s % 60, ms % 1000);
            INFO("Thanks. Share. Come again!");
        }
        if (cli.timing)
            fprintf(stderr, "timing: total %.3f s | read+decompress %.3f s | waiting for the GPU %.3f s | gather+compress+write %.3f s | "
                            "%llu bytes in, %llu records, %llu batches, %llu re-submitted\n",
                    rs.total_s, rs.read_s, rs.gpu_wait_s, rs.write_s, (unsigned long long)rs.bytes_in, (unsigned long long)rs.records,
                    (unsigned long long)rs.batches, (unsigned long long)rs.respeculated);
        if (cli.timing)
            fprintf(stderr, "timing: process %.3f s since main() | GPU setup (contexts, pinned slots, workspaces, match table) %.3f s\n",
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count(), rs.setup_s);
    } catch (const ShortReadPanic& e) {
        fprintf(stderr, "thread 'main' panicked: range end index out of range for slice (%s)\n", e.what());
        return EX_PANIC;
    } catch (const std::exception& e) {
        fprintf(stderr, "Error: %s\n", e.what());  // anyhow's report of main's Err
        return 1;
    }
    return 0;
}
