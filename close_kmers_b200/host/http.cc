// Request-head parsing and dispatch of KmerRequest2, without Boost.Regex: the three regular expressions of
// krequest2.cc:26-33 are matched by hand (each is anchored and has one way to match).
#include "http.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <stdexcept>

#include "../../include/ckm_server.h"

namespace ckm_http {

namespace {
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }
inline bool is_upper(char c) { return c >= 'A' && c <= 'Z'; }
}  // namespace

// request_regex = "^([A-Z]+) ([^?#]*)(\?([^#]*))?(#(.*))? HTTP/(\d+\.\d+)" under regex_match (whole line).
// The version is the " HTTP/<digits>.<digits>" that ends the line; what lies between the method and it splits at the
// first '?' or '#': path, then parameters up to the next '#', then the fragment.
bool Request::parse_request_line(std::string line) {
    const size_t cr = line.find('\r');
    if (cr != std::string::npos) line.erase(cr);
    size_t i = 0;
    while (i < line.size() && is_upper(line[i])) i++;
    if (i == 0 || i >= line.size() || line[i] != ' ') return false;
    const size_t e = line.size();
    size_t j = e;
    while (j > 0 && is_digit(line[j - 1])) j--;
    if (j == e || j == 0 || line[j - 1] != '.') return false;
    const size_t dot = j - 1;
    size_t k = dot;
    while (k > 0 && is_digit(line[k - 1])) k--;
    if (k == dot || k < 6 || line.compare(k - 6, 6, " HTTP/") != 0) return false;
    const size_t suffix = k - 6;  // the space before "HTTP/"
    if (suffix < i + 1) return false;
    const std::string middle = line.substr(i + 1, suffix - (i + 1));
    type = line.substr(0, i);
    version = line.substr(k);
    parameters_raw.clear();
    fragment.clear();
    const size_t sp = middle.find_first_of("?#");
    path = middle.substr(0, sp);
    if (sp != std::string::npos) {
        if (middle[sp] == '?') {
            const size_t h = middle.find('#', sp + 1);
            parameters_raw = middle.substr(sp + 1, h == std::string::npos ? std::string::npos : h - (sp + 1));
            if (h != std::string::npos) fragment = middle.substr(h + 1);
        } else {
            fragment = middle.substr(sp + 1);
        }
    }
    // boost::split(parts, parameters_raw_, is_any_of(";&")); parts without '=' are dropped; no URL decoding
    if (!parameters_raw.empty()) {
        size_t b = 0;
        for (;;) {
            const size_t t = parameters_raw.find_first_of(";&", b);
            const std::string part = parameters_raw.substr(b, t == std::string::npos ? std::string::npos : t - b);
            const size_t eq = part.find('=');
            if (eq != std::string::npos) parameters[part.substr(0, eq)] = part.substr(eq + 1);
            if (t == std::string::npos) break;
            b = t + 1;
        }
    }
    return true;
}

void Request::parse_header_line(std::string line) {
    const size_t cr = line.find('\r');
    if (cr != std::string::npos) line.erase(cr);
    size_t x = line.find(':');
    std::string k(line.substr(0, x));
    x = x == std::string::npos ? 0 : x + 1;
    while (x < line.size() && line[x] == ' ') x++;
    std::string v(line.substr(x));
    std::transform(k.begin(), k.end(), k.begin(), [](unsigned char c) { return (char)::tolower(c); });
    headers[k] = v;
}

const std::string &Request::param(const std::string &k) const {
    static const std::string empty;
    auto it = parameters.find(k);
    return it == parameters.end() ? empty : it->second;
}

static Decision respond(int code, const char *status, const std::string &body) {
    Decision d;
    d.kind = Decision::RESPOND;
    d.code = code;
    d.status = status;
    d.body = body;
    return d;
}

Decision decide(const Request &r) {
    auto te = r.headers.find("transfer-encoding");
    if (te != r.headers.end() && te->second == "chunked")
        return respond(501, "Chunked encoding not implemented", "Chunked encoding not implemented\n");
    Decision d;
    auto ex = r.headers.find("expect");
    const bool cont = ex != r.headers.end() && ex->second == "100-continue";
    if (r.type == "GET") {
        static const char genus_prefix[] = "/genus_lookup/";
        if (r.path == "/quit") {
            d.kind = Decision::GET_QUIT;
        } else if (r.path == "/version") {
            d.kind = Decision::GET_VERSION;
        } else if (r.path.compare(0, sizeof genus_prefix - 1, genus_prefix) == 0 && r.path.size() > sizeof genus_prefix - 1 &&
                   r.path.find('/', sizeof genus_prefix - 1) == std::string::npos) {  // ^/genus_lookup/([^/]+)$
            d.kind = Decision::GET_GENUS;
            d.genus = r.path.substr(sizeof genus_prefix - 1);
        } else {
            d = respond(404, "Not found", "path not found\n");
        }
    } else if (r.type == "POST") {
        auto cl = r.headers.find("content-length");
        if (cl == r.headers.end()) {
            d = respond(500, "Missing content length", "Missing content length header\n");
        } else {
            size_t len = 0;
            try {
                len = std::stoul(cl->second);
            } catch (std::exception &e) {  // the catch block of read_headers, krequest2.cc:217-223
                d = respond(500, "Failed", std::string("Caught exception ") + e.what() + "\n");
                d.send_continue = cont;
                return d;
            }
            std::string key, action(r.path);
            // mapping_path_regex = "^/mapping/([^/]+)(/(add|matrix|lookup))$"
            static const char mp[] = "/mapping/";
            if (r.path.compare(0, sizeof mp - 1, mp) == 0) {
                const size_t s = r.path.find('/', sizeof mp - 1);
                if (s != std::string::npos && s > sizeof mp - 1) {
                    const std::string tail = r.path.substr(s);
                    if (tail == "/add" || tail == "/matrix" || tail == "/lookup") {
                        key = r.path.substr(sizeof mp - 1, s - (sizeof mp - 1));
                        action = tail;
                    }
                }
            }
            if (action == "/add" || action == "/matrix" || action == "/lookup" || action == "/fq_lookup" || action == "/query") {
                d.kind = Decision::POST;
                d.key = key;
                d.action = action;
                d.content_length = len;
            } else {
                d = respond(404, "Not found", "path not found\n");
            }
        }
    } else {
        d.none = true;  // process_request does nothing for other methods; the connection just sits there
    }
    d.send_continue = cont;
    return d;
}

}  // namespace ckm_http

extern "C" void ckm_free_text(char *text);

extern "C" char *ckm_http_describe(const char *head, size_t n) {
    using namespace ckm_http;
    std::string text(head, n);
    std::istringstream is(text);
    std::string line;
    std::ostringstream os;
    Request r;
    if (!std::getline(is, line) || !r.parse_request_line(line)) {
        os << "decision=invalid\n";
    } else {
        while (std::getline(is, line)) {
            const size_t cr = line.find('\r');
            if (cr != std::string::npos) line.erase(cr);
            if (line.empty()) break;
            r.parse_header_line(line);
        }
        os << "type=" << r.type << "\npath=" << r.path << "\nparameters=" << r.parameters_raw << "\nfragment=" << r.fragment
           << "\nversion=" << r.version << "\n";
        for (auto &p : r.parameters) os << "param." << p.first << "=" << p.second << "\n";
        for (auto &h : r.headers) os << "header." << h.first << "=" << h.second << "\n";
        const Decision d = decide(r);
        if (d.send_continue) os << "continue=1\n";
        os << "decision=";
        if (d.none) os << "none";
        else if (d.kind == Decision::RESPOND) os << "respond " << d.code << " " << d.status;
        else if (d.kind == Decision::GET_QUIT) os << "get quit";
        else if (d.kind == Decision::GET_VERSION) os << "get version";
        else if (d.kind == Decision::GET_GENUS) os << "get genus_lookup " << d.genus;
        else os << "post " << d.action << " key=" << d.key << " length=" << d.content_length;
        os << "\n";
    }
    const std::string s = os.str();
    char *p = (char *)malloc(s.size() + 1);
    if (!p) return nullptr;
    memcpy(p, s.data(), s.size() + 1);
    return p;
}
