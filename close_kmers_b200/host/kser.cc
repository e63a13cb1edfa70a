// kser_b200: the kser front end (kser.cc, kserver.cc, krequest2.cc and the handler choreography of
// query_request.cc / add_request.cc / matrix_request.cc / lookup_request.cc / fq_process_request.cc) without Boost, with
// every body chunk computed by the GPU handlers of ckm_handlers.h.
//
// Same routes, same response bytes.  What differs is scheduling: the reference parses 1 MB socket buffers and posts each
// to a CPU worker; here a connection thread parses the body as it arrives and hands the GPU batches of `--batch-mb`
// residues (responses are per-sequence text in arrival order, so the byte stream does not depend on where batches are cut).
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <signal.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <unistd.h>
#include <dirent.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/ckm.h"
#include "../../include/ckm_handlers.h"
#include "../../include/ckm_server.h"
#include "http.h"
#include "lookup.h"
#include "seq_parser.h"

namespace {

using ckm_http::Decision;
using ckm_http::Request;

const uint32_t NO_FAMILY = 0xffffffffu;

struct Options {
    std::string listen_port, listen_port_file = "/dev/null", kmer_data, kmer_version, families_version, genus_mapping, families_file,
        pid_file;
    std::vector<std::string> families_nr;
    bool have_kmer_version = false, have_families_version = false, no_listen = false, daemonize = false, debug_http = false, help = false;
    int n_load_threads = 1, n_kmer_threads = 4;
    std::vector<int> devices;
    size_t batch_bytes = 16u << 20;
};

const char *USAGE =
    " [options] listen-port kmer-data-dir\n"
    "Allowed options:\n"
    "  -h [ --help ]                  show this help message\n"
    "  -l [ --listen-port ] arg       port to listen on. 0 means to choose a random port\n"
    "  -d [ --kmer-data-dir ] arg     kmer data directory\n"
    "  --listen-port-file arg         save the listen port to this file\n"
    "  --kmer-version arg             kmer data version string\n"
    "  --families-genus-mapping arg   genus name to taxid mapping file\n"
    "  --families-file arg            families file\n"
    "  --families-nr arg...           families NR data\n"
    "  --families-version arg         families data version string\n"
    "  --n-load-threads arg (=1)      sizes the NR load chunks exactly like the reference's thread pool\n"
    "  --n-kmer-threads arg (=4)      engines per device (one KmerGuts per worker thread in the reference): requests on\n"
    "                                 different connections overlap their GPU work and response formatting\n"
    "  --no-listen                    don't listen - just load data and quit\n"
    "  --daemonize                    run the service in the background\n"
    "  --pid-file arg                 write the process id to this file\n"
    "  --debug-http                   debug HTTP protocol\n"
    "  --device arg (=0)              CUDA device(s), comma separated: one engine per device\n"
    "  --batch-mb arg (=16)           residues handed to the GPU per batch\n"
    "accepted and ignored (CPU scheduling of the reference): --n-family-file-threads --n-inserter-threads\n"
    "  --peg-kmer-data --reserve-mapping --no-populate-mmap; not supported: --family-reps --kmer-family-distribution-file\n"
    "If the kmer data directory contains files families.dat and a\n"
    "directory families.nr it will be assumed that these files contain\n"
    "family data files and will be loaded at startup. A file VERSION\n"
    "will set the family version data to the contents of the first line of that file\n";

bool is_regular_file(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}
bool is_directory(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}
std::string first_line(const std::string &p) {
    std::ifstream f(p);
    std::string l;
    std::getline(f, l);
    return l;
}

// kser.cc:54-99 (boost::program_options: "--name value", "--name=value", two positionals)
bool parse_options(int argc, char **argv, Options &o, std::string &err) {
    std::vector<std::string> pos;
    auto needs_value = [](const std::string &n) {
        static const char *v[] = {"n-family-file-threads", "n-inserter-threads", "n-load-threads", "n-kmer-threads", "listen-port-file",
                                  "peg-kmer-data", "listen-port", "kmer-data-dir", "kmer-version", "families-genus-mapping",
                                  "families-file", "families-version", "family-reps", "kmer-family-distribution-file", "reserve-mapping",
                                  "pid-file", "device", "batch-mb"};
        for (auto x : v)
            if (n == x) return true;
        return false;
    };
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "-h") a = "--help";
        if (a == "-l") a = "--listen-port";
        if (a == "-d") a = "--kmer-data-dir";
        if (a.compare(0, 2, "--") != 0) {
            pos.push_back(a);
            continue;
        }
        std::string name = a.substr(2), value;
        bool has_value = false;
        const size_t eq = name.find('=');
        if (eq != std::string::npos) {
            value = name.substr(eq + 1);
            name.erase(eq);
            has_value = true;
        }
        if (name == "families-nr") {  // multitoken
            if (has_value) o.families_nr.push_back(value);
            while (i + 1 < argc && argv[i + 1][0] != '-') o.families_nr.push_back(argv[++i]);
            continue;
        }
        if (needs_value(name) && !has_value) {
            if (i + 1 >= argc) {
                err = "the required argument for option '--" + name + "' is missing";
                return false;
            }
            value = argv[++i];
        }
        try {
            if (name == "help") o.help = true;
            else if (name == "listen-port") o.listen_port = value;
            else if (name == "kmer-data-dir") o.kmer_data = value;
            else if (name == "listen-port-file") o.listen_port_file = value;
            else if (name == "kmer-version") o.kmer_version = value, o.have_kmer_version = true;
            else if (name == "families-version") o.families_version = value, o.have_families_version = true;
            else if (name == "families-genus-mapping") o.genus_mapping = value;
            else if (name == "families-file") o.families_file = value;
            else if (name == "n-load-threads") o.n_load_threads = std::max(1, std::stoi(value));
            else if (name == "n-kmer-threads") o.n_kmer_threads = std::max(1, std::stoi(value));
            else if (name == "no-listen") o.no_listen = true;
            else if (name == "daemonize") o.daemonize = true;
            else if (name == "debug-http") o.debug_http = true;
            else if (name == "pid-file") o.pid_file = value;
            else if (name == "batch-mb") o.batch_bytes = (size_t)std::max(1, std::stoi(value)) << 20;
            else if (name == "device") {
                std::stringstream ss(value);
                std::string tok;
                while (std::getline(ss, tok, ',')) o.devices.push_back(std::stoi(tok));
            } else if (name == "family-reps" || name == "kmer-family-distribution-file") {
                std::cerr << "Warning: --" << name << " is not supported and is ignored\n";
            } else if (name == "n-family-file-threads" || name == "n-inserter-threads" ||
                       name == "peg-kmer-data" || name == "reserve-mapping" || name == "no-populate-mmap") {
            } else {
                err = "unrecognised option '--" + name + "'";
                return false;
            }
        } catch (std::exception &) {
            err = "the argument ('" + value + "') for option '--" + name + "' is invalid";
            return false;
        }
    }
    if (o.listen_port.empty() && !pos.empty()) {
        o.listen_port = pos.front();
        pos.erase(pos.begin());
    }
    if (o.kmer_data.empty() && !pos.empty()) {
        o.kmer_data = pos.front();
        pos.erase(pos.begin());
    }
    if (o.help) return true;
    if (o.listen_port.empty()) err = "the option '--listen-port' is required but missing";
    else if (o.kmer_data.empty()) err = "the option '--kmer-data-dir' is required but missing";
    return err.empty();
}

// kser.cc:110-180: family and version files found in the data directory become options (the command line wins)
void discover_data_dir(Options &o) {
    const std::string d = o.kmer_data + "/";
    std::string kversion("unknown"), fversion("unknown");
    if (is_regular_file(d + "VERSION")) {
        kversion = first_line(d + "VERSION");
        fversion = kversion;
    }
    if (is_regular_file(d + "families.version")) fversion = first_line(d + "families.version");
    if (!o.have_kmer_version) o.kmer_version = kversion, o.have_kmer_version = true;
    if (!o.have_families_version) o.families_version = fversion, o.have_families_version = true;
    if (o.genus_mapping.empty() && is_regular_file(d + "families.genus_map")) o.genus_mapping = d + "families.genus_map";
    if (o.families_file.empty() && is_regular_file(d + "families.dat")) o.families_file = d + "families.dat";
    if (o.families_nr.empty() && is_directory(d + "families.nr")) {
        if (DIR *dir = opendir((d + "families.nr").c_str())) {
            while (dirent *e = readdir(dir)) {
                const std::string p = d + "families.nr/" + e->d_name;
                if (is_regular_file(p)) o.families_nr.push_back(p);
            }
            closedir(dir);
        }
        std::sort(o.families_nr.begin(), o.families_nr.end());  // directory_iterator order is unspecified
    }
}

struct Engine {
    ckm_ctx *ctx = nullptr;
    std::mutex busy;
};

struct Mapping {  // one KmerPegMapping: peg ids on the host, postings on engine 0 under `post_key`
    ckm_mapping *ids = nullptr;
    uint32_t post_key = 0;
};

constexpr int kMaxConnections = 256;                 // concurrent connections (a thread each)
constexpr int kSocketTimeoutSeconds = 600;           // SO_RCVTIMEO / SO_SNDTIMEO of accepted sockets
constexpr size_t kMaxMatrixBody = (size_t)4 << 30;   // /matrix buffers the whole body

struct Server {
    Options opt;
    std::vector<std::unique_ptr<Engine>> engines;
    std::mutex map_mutex;
    std::map<std::string, Mapping> mappings;
    ckm_lookup::FamilyInfo fams;  // family_data_ + genus map
    std::vector<uint32_t> peg_to_family;  // by encoded id of the root mapping
    bool family_mode = false;
    std::atomic<bool> stopping{false};
    std::atomic<int> active{0};
    int wake_pipe[2] = {-1, -1};
    std::atomic<size_t> next_engine{0};

    Mapping &mapping_for(const std::string &key) {  // krequest2.cc:447-456
        std::lock_guard<std::mutex> g(map_mutex);
        auto it = mappings.find(key);
        if (it != mappings.end()) return it->second;
        Mapping m;
        m.ids = ckm_mapping_new();
        m.post_key = (uint32_t)mappings.size();
        return mappings.emplace(key, m).first->second;
    }
    // any free engine, else wait for the next one in turn
    Engine &lease(std::unique_lock<std::mutex> &lock, bool first_only) {
        if (!first_only) {
            for (auto &e : engines) {
                lock = std::unique_lock<std::mutex>(e->busy, std::try_to_lock);
                if (lock.owns_lock()) return *e;
            }
        }
        Engine &e = first_only ? *engines[0] : *engines[next_engine++ % engines.size()];
        lock = std::unique_lock<std::mutex>(e.busy);
        return e;
    }
};

Server *g_server = nullptr;

void on_signal(int signo) {
    if (g_server && g_server->wake_pipe[1] >= 0) {
        const char c = (char)signo;
        ssize_t r = write(g_server->wake_pipe[1], &c, 1);
        (void)r;
    }
}

std::vector<std::string> split_tabs(const std::string &line) {
    std::vector<std::string> cols;
    size_t b = 0;
    for (;;) {
        const size_t t = line.find('\t', b);
        cols.push_back(line.substr(b, t == std::string::npos ? std::string::npos : t - b));
        if (t == std::string::npos) break;
        b = t + 1;
    }
    return cols;
}

// KmerPegMapping::load_genus_map, kmer.cc:338-355
void load_genus_map(Server &s, const std::string &file) {
    std::ifstream gf(file);
    if (gf.fail()) {
        std::cerr << "Error opening gnus file " << file << "\n";
        exit(1);
    }
    std::string line;
    while (std::getline(gf, line)) {
        auto cols = split_tabs(line);
        if (cols.size() >= 2) s.fams.genus_map[cols[0]] = cols[1];
    }
}

// KmerPegMapping::load_families, kmer.cc:375-494, read by one thread (family ids in order of first appearance).
// Columns: 0 global family ("GF..."), 3 peg, 4 sequence length, 5 function, 7 genus, 8 local family number.
void load_families(Server &s, const std::string &file) {
    std::ifstream f(file);
    if (f.fail()) {
        std::cerr << "Failure opening families file " << file << "\n";
        exit(1);
    }
    Mapping &root = s.mapping_for("");
    std::map<std::pair<std::string, std::string>, uint32_t> family_key_to_id;
    std::map<std::string, bool> warned;
    const std::string zeros("00000000");
    std::string line;
    while (std::getline(f, line)) {
        auto cols = split_tabs(line);
        if (cols.size() < 9 || cols[0].size() < 2) continue;
        std::string pgf("PGF_");
        pgf += cols[0].substr(2);
        std::string plf("PLF_");
        unsigned long genus_id = 0;
        auto mapped = s.fams.genus_map.find(cols[7]);
        if (mapped == s.fams.genus_map.end()) {
            if (!warned[cols[7]]) {
                std::cerr << "Cannot map genus '" << cols[7] << "' in " << file << "\n";
                warned[cols[7]] = true;
            }
            plf += cols[7];
        } else {
            plf += mapped->second;
            try {
                genus_id = std::stoul(mapped->second);
            } catch (std::exception &) {
            }
        }
        plf += "_";
        plf += zeros.substr(0, cols[8].size() < 8 ? 8 - cols[8].size() : 0);
        plf += cols[8];
        const uint32_t id = ckm_mapping_assign_new_id(root.ids, cols[3].c_str());
        unsigned long seqlen = 0;
        try {
            seqlen = std::stoul(cols[4]);
        } catch (std::exception &) {
        }
        const auto fkey = std::make_pair(pgf, plf);
        uint32_t fam_id;
        auto it = family_key_to_id.find(fkey);
        if (it == family_key_to_id.end()) {
            fam_id = (uint32_t)s.fams.data.size();
            family_key_to_id[fkey] = fam_id;
            s.fams.data.push_back({pgf, plf, cols[5], genus_id, seqlen, 1});
        } else {
            fam_id = it->second;
            s.fams.data[fam_id].total_size += seqlen;
            s.fams.data[fam_id].count++;
        }
        if (s.peg_to_family.size() <= id) s.peg_to_family.resize((size_t)id + 1, NO_FAMILY);
        s.peg_to_family[id] = fam_id;
    }
}

int fail_ckm(const char *what) {
    std::cerr << what << ": " << ckm_last_error() << "\n";
    return 1;
}

// NRLoader::load_families + thread_load (nr_loader.cc:46-202) for one file: the FASTA records are cut into chunks of
// >= max_size residues; in family mode a protein without a family ends its chunk (the rest of the chunk is dropped,
// nr_loader.cc:154-160), which is applied here so that whole GPU batches can be sent without further bookkeeping.
int load_nr_file(Server &s, const std::string &file, size_t n_files) {
    struct stat st;
    if (stat(file.c_str(), &st) != 0) {
        std::cerr << "cannot stat " << file << "\n";
        return 1;
    }
    size_t max_size = (size_t)st.st_size / (size_t)s.opt.n_load_threads / (size_t)std::ceil(10.0 / (float)n_files);
    if (max_size < 1000000) max_size = 1000000;
    std::cerr << "Begin load of " << file << "\ntp size=" << s.opt.n_load_threads << " max_size_=" << max_size << "\n";
    FILE *fp = fopen(file.c_str(), "rb");
    if (!fp) {
        std::cerr << "cannot open " << file << "\n";
        return 1;
    }
    Mapping &root = s.mapping_for("");
    ckm_ctx *ctx = s.engines[0]->ctx;
    ckm_seq_parser *parser = ckm_seq_parser_new(CKM_FORMAT_FASTA);
    std::vector<uint32_t> ids;       // family id (family mode) or encoded peg id per kept sequence
    std::string residues;
    std::vector<uint64_t> offsets{0};
    size_t cur_size = 0;
    bool chunk_dropped = false;
    uint64_t n_seqs = 0, n_dropped = 0;
    int rc = 0;
    auto flush = [&]() {
        if (ids.empty()) return 0;
        int r = s.family_mode ? ckm_family_nr_add(ctx, ids.data(), residues.data(), offsets.data(), (uint32_t)ids.size())
                              : ckm_postings_add(ctx, ids.data(), residues.data(), offsets.data(), (uint32_t)ids.size());
        ids.clear();
        residues.clear();
        offsets.assign(1, 0);
        return r;
    };
    auto consume = [&](const ckm_seq_batch_t &b) {
        for (uint32_t i = 0; i < b.n && !rc; i++) {
            const size_t len = (size_t)(b.offsets[i + 1] - b.offsets[i]);
            const uint32_t enc = ckm_mapping_encode_id(root.ids, b.ids[i]);
            n_seqs++;
            if (!chunk_dropped) {
                uint32_t tag = enc;
                if (s.family_mode) {
                    tag = enc < s.peg_to_family.size() ? s.peg_to_family[enc] : NO_FAMILY;
                    if (tag == NO_FAMILY) {
                        std::cerr << "NO FAM FOR id='" << b.ids[i] << "' enc_id='" << enc << "' dec='" << b.ids[i] << "'\n";
                        chunk_dropped = true;
                    }
                }
                if (!chunk_dropped) {
                    ids.push_back(tag);
                    residues.append(b.residues + b.offsets[i], len);
                    offsets.push_back(residues.size());
                }
            }
            if (chunk_dropped) n_dropped++;
            cur_size += len;
            if (cur_size >= max_size) {  // on_parsed_seq, nr_loader.cc:94-108: the chunk ends here
                cur_size = 0;
                chunk_dropped = false;
                if (residues.size() >= s.opt.batch_bytes) rc = flush();
            }
        }
    };
    std::vector<char> buf(8u << 20);
    ckm_seq_batch_t b;
    size_t got;
    while (!rc && (got = fread(buf.data(), 1, buf.size(), fp)) > 0) {
        ckm_seq_parser_feed(parser, buf.data(), got);
        ckm_seq_parser_take(parser, &b);
        consume(b);
    }
    fclose(fp);
    // parser.parse(inp) ends with parse_complete, and load_families calls it once more (nr_loader.cc:66-67): the last
    // record, then an empty one
    ckm_seq_parser_complete(parser);
    ckm_seq_parser_complete(parser);
    ckm_seq_parser_take(parser, &b);
    consume(b);
    if (!rc) rc = flush();
    ckm_seq_parser_free(parser);
    if (rc) return fail_ckm("loading families NR");
    std::cerr << file << " loader completed: " << n_seqs << " sequences, " << n_dropped << " dropped with their chunk tail\n";
    std::cout << "load complete on " << file << "\n";
    return 0;
}

int install_families(Server &s) {
    std::vector<const char *> pgf, plf, fn;
    for (auto &f : s.fams.data) {
        pgf.push_back(f.pgf.c_str());
        plf.push_back(f.plf.c_str());
        fn.push_back(f.function.c_str());
    }
    ckm_ctx *c0 = s.engines[0]->ctx;
    uint64_t nk = 0, ne = 0;
    if (ckm_family_nr_finish(c0, (uint32_t)pgf.size(), pgf.data(), plf.data(), fn.data(), &nk, &ne)) return fail_ckm("family table");
    std::cerr << "family table: " << nk << " k-mers, " << ne << " entries, " << pgf.size() << " families\n";
    if (s.engines.size() > 1) {  // replicate to the other devices
        std::vector<uint64_t> kmers(nk), off(nk + 1);
        std::vector<uint32_t> ids(ne + 1);
        if (ckm_family_export(c0, nk, ne, kmers.data(), off.data(), ids.data())) return fail_ckm("family export");
        for (size_t e = 1; e < s.engines.size(); e++)
            if (ckm_family_load(s.engines[e]->ctx, nk, kmers.data(), off.data(), ids.data(), (uint32_t)pgf.size(), pgf.data(), plf.data(),
                                fn.data()))
                return fail_ckm("family load");
    }
    return 0;
}

// ---- one connection ------------------------------------------------------------------------------------------------

struct Conn {
    int fd;
    std::vector<char> buf;
    size_t pos = 0, end = 0;
    explicit Conn(int f) : fd(f), buf(1u << 20) {}  // the reference's streambuf is 1 MB too (krequest2.cc:41)
    bool fill() {
        if (pos == end) pos = end = 0;
        if (end == buf.size()) {
            if (pos == 0) return false;  // a line longer than the buffer
            memmove(buf.data(), buf.data() + pos, end - pos);
            end -= pos;
            pos = 0;
        }
        ssize_t r;
        do r = recv(fd, buf.data() + end, buf.size() - end, 0);
        while (r < 0 && errno == EINTR);
        if (r <= 0) return false;
        end += (size_t)r;
        return true;
    }
    bool read_line(std::string &line) {
        size_t scanned = pos;
        for (;;) {
            const char *nl = (const char *)memchr(buf.data() + scanned, '\n', end - scanned);
            if (nl) {
                line.assign((const char *)buf.data() + pos, nl);
                pos = (size_t)(nl - buf.data()) + 1;
                return true;
            }
            const size_t had = end - pos;
            if (!fill()) return false;
            scanned = pos + had;
        }
    }
    size_t read_some(const char *&p, size_t max) {
        if (pos == end && !fill()) return 0;
        const size_t n = std::min(max, end - pos);
        p = buf.data() + pos;
        pos += n;
        return n;
    }
    bool write_all(const char *p, size_t n) {
        while (n) {
            ssize_t w = send(fd, p, n, MSG_NOSIGNAL);
            if (w < 0 && errno == EINTR) continue;
            if (w <= 0) return false;
            p += w;
            n -= (size_t)w;
        }
        return true;
    }
    bool write_all(const std::string &s) { return write_all(s.data(), s.size()); }
};

// KmerRequest2::write_header + respond, krequest2.cc:488-518
std::string header_text(const Request &r, int code, const std::string &status) {
    return "HTTP/" + r.version + " " + std::to_string(code) + " " + status + "\nContent-type: text/plain\n";
}
void respond(Conn &c, const Request &r, int code, const std::string &status, const std::string &body) {
    c.write_all(header_text(r, code, status) + "Content-length: " + std::to_string(body.size()) + "\n\n" + body);
}

// KmerGuts::set_parameters (kguts.cc:244-268) from the request's parameters, before every batch like the handlers
int apply_parameters(ckm_ctx *ctx, const Request &r) {
    int v[4] = {0, 5, 0, 200};
    static const char *names[4] = {"order_constraint", "min_hits", "min_weighted_hits", "max_gap"};
    for (int k = 0; k < 4; k++) {
        auto it = r.parameters.find(names[k]);
        if (it == r.parameters.end()) continue;
        try {
            v[k] = std::stoi(it->second);
        } catch (const std::invalid_argument &) {
            std::cerr << "Warning: invalid integer value '" << it->second << "' for parameter " << names[k] << "\n";
        } catch (const std::out_of_range &) {
        }
    }
    return ckm_set_params(ctx, v[0], v[1], v[2], v[3]);
}

int int_param(const Request &r, const char *name) {  // try { x = std::stoi(parameters()[name]); } catch (...) {}
    try {
        return std::stoi(r.param(name));
    } catch (...) {
        return 0;
    }
}

struct Inflater {  // a gzip body on /fq_lookup (fq_process_request.cc:66-97, zlib_support)
    z_stream zs;
    bool active = false, failed = false;
    std::vector<char> out;
    Inflater() : out(4u << 20) { memset(&zs, 0, sizeof zs); }
    ~Inflater() {
        if (active) inflateEnd(&zs);
    }
    bool start() {
        active = inflateInit2(&zs, 15 + 16) == Z_OK;
        return active;
    }
    template <class Sink>
    void feed(const char *p, size_t n, Sink sink) {
        zs.next_in = (Bytef *)p;
        zs.avail_in = (uInt)n;
        while (zs.avail_in && !failed) {
            zs.next_out = (Bytef *)out.data();
            zs.avail_out = (uInt)out.size();
            const int rc = inflate(&zs, Z_NO_FLUSH);
            sink(out.data(), out.size() - zs.avail_out);
            if (rc == Z_STREAM_END) {
                if (zs.avail_in) inflateReset(&zs);  // next member
            } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
                failed = true;
            } else if (rc == Z_BUF_ERROR && zs.avail_out) {
                break;
            }
        }
    }
};

// One parsed batch on its way from the connection thread (socket + parser) to the request's worker (GPU + text + send).
struct Job {
    std::vector<std::string> ids;
    std::string residues;
    std::vector<uint64_t> offsets;
};

struct JobQueue {  // depth 2: the connection thread parses batch k+1 while batch k is computed, formatted and sent
    std::mutex m;
    std::condition_variable cv;
    std::vector<std::unique_ptr<Job>> q;
    bool closed = false, failed = false;
    bool push(std::unique_ptr<Job> j) {
        std::unique_lock<std::mutex> l(m);
        cv.wait(l, [&] { return q.size() < 2 || failed; });
        if (failed) return false;
        q.push_back(std::move(j));
        cv.notify_all();
        return true;
    }
    std::unique_ptr<Job> pop() {
        std::unique_lock<std::mutex> l(m);
        cv.wait(l, [&] { return !q.empty() || closed; });
        if (q.empty()) return nullptr;
        std::unique_ptr<Job> j = std::move(q.front());
        q.erase(q.begin());
        cv.notify_all();
        return j;
    }
    void close() {
        std::lock_guard<std::mutex> l(m);
        closed = true;
        cv.notify_all();
    }
    void fail() {
        std::lock_guard<std::mutex> l(m);
        failed = true;
        cv.notify_all();
    }
};

void handle_post(Server &s, Conn &c, const Request &r, const Decision &d) {
    const std::string &action = d.action;
    const bool is_fq = action == "/fq_lookup", is_matrix = action == "/matrix", is_add = action == "/add", is_lookup = action == "/lookup";
    if (is_fq && d.content_length == 0) {  // fq_process_request.cc:43-47
        respond(c, r, 200, "OK", "data done\n");
        return;
    }
    Mapping &mapping = s.mapping_for(d.key);
    bool header_written = false;
    if (is_add) {  // krequest2.cc:458-466: the header goes out before any data is read
        if (!c.write_all("HTTP/" + r.version + " 200 OK\nContent-type: text/plain\n\n")) return;
        header_written = true;
    }
    const int silent = is_add ? [&] {
        try {
            return std::stoi(r.param("silent"));
        } catch (const std::invalid_argument &) {
            return 0;
        } catch (const std::out_of_range &) {
            return 0;
        }
    }() : 0;
    ckm_lookup_options_t lopt;
    memset(&lopt, 0, sizeof lopt);
    if (is_lookup) lopt = ckm_lookup::options_from(r, s.fams, s.family_mode);

    // ---- worker: one batch at a time, in arrival order ----
    JobQueue jobs;
    auto work = [&]() {
        while (std::unique_ptr<Job> j = jobs.pop()) {
            std::vector<const char *> idp;
            idp.reserve(j->ids.size());
            for (const auto &id : j->ids) idp.push_back(id.c_str());
            const uint32_t n = (uint32_t)j->ids.size();
            std::string out;
            if (!header_written && !is_matrix) {
                out = header_text(r, 200, "OK") + "\n";
                header_written = true;
            }
            char *text = nullptr;
            int rc = 0;
            {
                std::unique_lock<std::mutex> lock;
                Engine &e = s.lease(lock, is_add || is_matrix || (is_lookup && !s.family_mode));
                if (!is_fq) rc = apply_parameters(e.ctx, r);  // fq_process_request.cc never calls set_parameters
                if (rc) {
                } else if (action == "/query") {
                    rc = ckm_query_text(e.ctx, idp.data(), j->residues.data(), j->offsets.data(), n, int_param(r, "details"),
                                        int_param(r, "find_best_call"), &text);
                } else if (is_add) {
                    rc = ckm_postings_select(e.ctx, mapping.post_key);
                    if (!rc) rc = ckm_add_text(e.ctx, mapping.ids, idp.data(), j->residues.data(), j->offsets.data(), n, silent, &text);
                } else if (is_matrix) {
                    rc = ckm_postings_select(e.ctx, mapping.post_key);
                    if (!rc) rc = ckm_matrix_text(e.ctx, mapping.ids, idp.data(), j->residues.data(), j->offsets.data(), n, &text);
                    out = "HTTP/1.1 200 OK\nContent-type: text/plain\n\n";  // matrix_request.cc:168-170
                } else if (is_fq) {
                    rc = ckm_fq_text(e.ctx, idp.data(), j->residues.data(), j->offsets.data(), n, &text);
                } else {
                    if (!s.family_mode) rc = ckm_postings_select(e.ctx, mapping.post_key);
                    if (!rc)
                        rc = ckm_lookup_text(e.ctx, mapping.ids, s.fams.flat.data(), (uint32_t)s.fams.flat.size(), &lopt, idp.data(),
                                             j->residues.data(), j->offsets.data(), n, &text);
                }
            }
            bool ok = true;
            if (rc) {
                std::cerr << "ERROR in " << action << ": " << ckm_last_error() << "\n";
                if (!header_written || is_matrix) respond(c, r, 500, "Failed", std::string("Caught exception ") + ckm_last_error() + "\n");
                ok = false;
            } else {
                ok = c.write_all(out) && (!text || c.write_all(text, strlen(text)));  // no second copy of a large response
            }
            ckm_free_text(text);
            if (!ok) {
                jobs.fail();
                return;
            }
        }
    };
    if (is_matrix && d.content_length > kMaxMatrixBody) {  // /matrix holds the whole request in memory (as the reference does): bounded
        respond(c, r, 500, "Failed", "Caught exception request body too large\n");
        return;
    }
    std::thread worker(work);
    struct Join {  // whatever leaves this function -- a bad_alloc in the parser included -- closes the queue and joins the worker
        JobQueue &q;
        std::thread &t;
        ~Join() {
            q.close();
            if (t.joinable()) t.join();
        }
    } join_guard{jobs, worker};

    // ---- this thread: socket -> (inflate) -> parser -> batches ----
    ckm_seq_parser *parser = ckm_seq_parser_new(is_fq ? CKM_FORMAT_FASTQ : CKM_FORMAT_FASTA);
    struct FreeParser {
        ckm_seq_parser *p;
        ~FreeParser() { ckm_seq_parser_free(p); }
    } parser_guard{parser};
    Inflater gz;
    size_t remaining = d.content_length;
    bool first_bytes = true, finished = false, ok = true;
    while (ok && !finished) {
        if (remaining == 0) {
            finished = true;
        } else {
            const char *p;
            const size_t n = c.read_some(p, remaining);
            if (n == 0) {
                finished = true;  // eof before content-length bytes
            } else {
                if (first_bytes && is_fq && n >= 2 && (unsigned char)p[0] == 0x1f && (unsigned char)p[1] == 0x8b) gz.start();
                first_bytes = false;
                if (gz.active) gz.feed(p, n, [&](const char *q, size_t m) { ckm_seq_parser_feed(parser, q, m); });
                else ckm_seq_parser_feed(parser, p, n);
                remaining -= n;
                if (remaining == 0) finished = true;
            }
        }
        if (finished) ckm_seq_parser_complete(parser);
        if (!finished && (is_matrix || ckm_seq_parser_pending(parser) < s.opt.batch_bytes)) continue;
        std::unique_ptr<Job> j(new Job());
        parser->take_owned(j->ids, j->residues, j->offsets);
        ok = jobs.push(std::move(j));
    }
    jobs.close();
    worker.join();
}

void handle_connection(Server *sp, int fd) {
    Server &s = *sp;
    Conn c(fd);
    std::string line;
    Request r;
    bool quit = false;
    do {
        if (!c.read_line(line)) break;
        {
            std::string shown = line;
            const size_t cr = shown.find('\r');
            if (cr != std::string::npos) shown.erase(cr);
            if (s.opt.debug_http) std::cerr << "Request: " << shown << "\n";
            if (!r.parse_request_line(line)) {
                for (char &ch : shown)  // stringPurifier (krequest2.cc:76-85; defined there, never called): keep the log printable
                    if ((unsigned char)ch < 32 || (unsigned char)ch > 127) ch = ' ';
                if (shown.size() > 200) shown = shown.substr(0, 200) + "...";
                std::cerr << "Invalid request '" << shown << "'\n";
                break;
            }
        }
        bool finished = false;
        while (!finished && c.read_line(line)) {
            const size_t cr = line.find('\r');
            if (cr != std::string::npos) line.erase(cr);
            if (line.empty()) finished = true;
            else r.parse_header_line(line);
        }
        if (!finished) break;
        if (s.opt.debug_http) {
            std::cerr << "Headers:\n";
            for (auto &h : r.headers) std::cerr << h.first << ": " << h.second << "\n";
        }
        const Decision d = ckm_http::decide(r);
        if (d.send_continue) c.write_all("HTTP/" + r.version + " 100 Continue\n\n");
        if (d.none) break;
        switch (d.kind) {
        case Decision::RESPOND:
            respond(c, r, d.code, d.status, d.body);
            break;
        case Decision::GET_QUIT:
            respond(c, r, 200, "OK", "OK, quitting\n");
            quit = true;
            break;
        case Decision::GET_VERSION: {  // krequest2.cc:274-289
            std::string body;
            if (s.opt.have_kmer_version) body += "kmer\t" + s.opt.kmer_version + "\n";
            if (s.opt.have_families_version) body += "families\t" + s.opt.families_version + "\n";
            body += std::string("family-mode\t") + (s.family_mode ? "1" : "0") + "\n";
            respond(c, r, 200, "OK", body);
            break;
        }
        case Decision::GET_GENUS: {  // krequest2.cc:290-312
            auto hit = s.fams.genus_map.find(d.genus);
            if (hit == s.fams.genus_map.end()) respond(c, r, 404, "Not Found", "genus not found\n");
            else respond(c, r, 200, "OK", hit->second + "\n");
            break;
        }
        case Decision::POST:
            try {
                handle_post(s, c, r, d);
            } catch (std::exception &e) {
                std::cerr << "Caught exception " << e.what() << "\n";
            }
            break;
        }
    } while (false);
    shutdown(fd, SHUT_RDWR);
    close(fd);
    if (quit) {
        std::cerr << "stopping io service\n";
        on_signal(0);
    }
    s.active--;
}

}  // namespace

extern "C" int ckm_kser_main(int argc, char **argv) {
    Server s;
    std::string err;
    if (!parse_options(argc, argv, s.opt, err)) {
        std::cerr << "Invalid command line: " << err << "\nUsage: " << argv[0] << USAGE;
        return 1;
    }
    if (s.opt.help) {
        std::cout << "Usage: " << argv[0] << USAGE;
        return 1;
    }
    discover_data_dir(s.opt);
    if (s.opt.devices.empty()) s.opt.devices.push_back(0);
    s.family_mode = !s.opt.families_file.empty();  // kser.cc:280

    if (s.opt.daemonize) {  // kser.cc:209-235
        pid_t child = fork();
        if (child < 0) {
            std::cerr << "fork failed: " << strerror(errno) << "\n";
            return 1;
        }
        if (child > 0) {
            if (!s.opt.pid_file.empty()) std::ofstream(s.opt.pid_file) << child << "\n";
            return 0;
        }
        if (setsid() < 0) {
            std::cerr << "setsid failed: " << strerror(errno) << "\n";
            return 1;
        }
    } else if (!s.opt.pid_file.empty()) {
        std::ofstream(s.opt.pid_file) << getpid() << "\n";
    }

    for (int dev : s.opt.devices) {
        std::unique_ptr<Engine> e(new Engine());
        if (ckm_open(s.opt.kmer_data.c_str(), dev, &e->ctx)) return fail_ckm("ckm_open");
        s.engines.push_back(std::move(e));
    }
    s.mapping_for("");  // the root mapping, kserver.cc:30-33

    // KmerRequestServer::KmerRequestServer, kserver.cc:35-127
    if (!s.opt.genus_mapping.empty()) load_genus_map(s, s.opt.genus_mapping);
    if (s.family_mode) {
        std::cerr << "Loading (immediate) families from " << s.opt.families_file << "...\n";
        load_families(s, s.opt.families_file);
        std::cerr << "Loading families from " << s.opt.families_file << "... done\n";
    }
    if (ckm_family_nr_begin(s.engines[0]->ctx)) return fail_ckm("family table");
    for (auto &f : s.opt.families_nr) {
        std::cerr << "Queue load NR file " << f << "\n";
        if (load_nr_file(s, f, s.opt.families_nr.size())) return 1;
    }
    if (install_families(s)) return 1;
    s.fams.flatten();
    // the worker engines: clones share the tables of their device's first engine (threadpool.cc:33)
    {
        const size_t n_dev = s.engines.size();
        for (int k = 1; k < s.opt.n_kmer_threads; k++)
            for (size_t dv = 0; dv < n_dev; dv++) {
                std::unique_ptr<Engine> e(new Engine());
                if (ckm_clone(s.engines[dv]->ctx, &e->ctx)) return fail_ckm("ckm_clone");
                s.engines.push_back(std::move(e));
            }
    }
    if (s.opt.no_listen) {
        std::cerr << "Quitting due to --no-listen being set\n";
        return 0;
    }

    // KmerRequestServer::startup, kserver.cc:132-158
    g_server = &s;
    if (pipe(s.wake_pipe) != 0) return 1;
    struct sigaction sa;
    memset(&sa, 0, sizeof sa);
    sa.sa_handler = on_signal;
    sigaction(SIGINT, &sa, nullptr);
    sigaction(SIGTERM, &sa, nullptr);
    sigaction(SIGQUIT, &sa, nullptr);
    signal(SIGPIPE, SIG_IGN);
    const int lfd = socket(AF_INET, SOCK_STREAM, 0);
    int one = 1;
    setsockopt(lfd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    sockaddr_in addr;
    memset(&addr, 0, sizeof addr);
    addr.sin_family = AF_INET;
    addr.sin_addr.s_addr = htonl(INADDR_ANY);
    addr.sin_port = htons((uint16_t)atoi(s.opt.listen_port.c_str()));
    if (lfd < 0 || bind(lfd, (sockaddr *)&addr, sizeof addr) != 0 || listen(lfd, 128) != 0) {
        std::cerr << "cannot listen on port " << s.opt.listen_port << ": " << strerror(errno) << "\n";
        return 1;
    }
    socklen_t alen = sizeof addr;
    getsockname(lfd, (sockaddr *)&addr, &alen);
    std::cout << "Listening on 0.0.0.0:" << ntohs(addr.sin_port) << "\n" << std::flush;
    if (!s.opt.listen_port_file.empty()) std::ofstream(s.opt.listen_port_file) << ntohs(addr.sin_port) << "\n";

    for (;;) {
        pollfd pf[2] = {{lfd, POLLIN, 0}, {s.wake_pipe[0], POLLIN, 0}};
        if (poll(pf, 2, -1) < 0) {
            if (errno == EINTR) continue;
            break;
        }
        if (pf[1].revents) {
            char signo = 0;
            ssize_t rr = read(s.wake_pipe[0], &signo, 1);
            (void)rr;
            if (signo) std::cout << "Exiting with signal " << (int)signo << "\n";
            break;
        }
        if (pf[0].revents & POLLIN) {
            const int fd = accept(lfd, nullptr, nullptr);
            if (fd < 0) continue;
            if (s.active.load() >= kMaxConnections) {  // one thread per connection: bounded
                static const char busy[] = "HTTP/1.1 503 Service Unavailable\nContent-type: text/plain\n\nToo many connections\n";
                ssize_t w = write(fd, busy, sizeof busy - 1);
                (void)w;
                close(fd);
                continue;
            }
            setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
            timeval tmo = {kSocketTimeoutSeconds, 0};  // a client that stops sending or reading does not hold a thread for ever
            setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tmo, sizeof tmo);
            setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tmo, sizeof tmo);
            s.active++;
            try {
                std::thread(handle_connection, &s, fd).detach();
            } catch (const std::system_error &) {
                s.active--;
                close(fd);
            }
        }
    }
    close(lfd);
    s.stopping = true;
    for (int i = 0; i < 3000 && s.active.load() > 0; i++) usleep(10000);  // let the requests in flight finish
    if (s.active.load() == 0) {
        for (size_t e = s.engines.size(); e-- > 0;) ckm_close(s.engines[e]->ctx);  // clones first
        for (auto &m : s.mappings) ckm_mapping_free(m.second.ids);
    }
    else {
        // connection threads still hold pointers into `s`, which lives on the caller's stack: do not return underneath them
        std::cout << std::flush;
        std::cerr << "kser_b200: " << s.active.load() << " connection(s) still busy after 30 s; exiting\n" << std::flush;
        _exit(0);
    }
    g_server = nullptr;
    return 0;
}
