// Request head of KmerRequest2 (krequest2.cc): request line, parameters, headers, and what process_request does next.
// Internal to the host layer; the C test hook is ckm_http_describe (ckm_server.h).
#ifndef CKM_HOST_HTTP_H
#define CKM_HOST_HTTP_H
#include <cstddef>
#include <map>
#include <string>

namespace ckm_http {

struct Request {
    std::string type, path, parameters_raw, fragment, version;  // request_regex groups 1, 2, 4, 6, 7
    std::map<std::string, std::string> parameters, headers;

    // read_initial_line (krequest2.cc:87-159): false when the line does not match request_regex
    bool parse_request_line(std::string line);
    // one header line of read_headers (krequest2.cc:171-192); the caller stops at the empty line
    void parse_header_line(std::string line);
    // operator[] semantics of owner_->parameters()["x"]
    const std::string &param(const std::string &k) const;
};

struct Decision {
    enum Kind { RESPOND, GET_QUIT, GET_VERSION, GET_GENUS, POST } kind = RESPOND;
    bool send_continue = false;        // Expect: 100-continue (krequest2.cc:253-261)
    int code = 0;                      // RESPOND
    std::string status, body;          // RESPOND
    std::string genus;                 // GET_GENUS
    std::string key, action;           // POST: mapping key ("" = root) and "/add" ...
    size_t content_length = 0;         // POST
    bool none = false;                 // a request type the reference ignores (neither GET nor POST)
};

// read_headers' chunked check + process_request's dispatch (krequest2.cc:208-213, 247-486), up to the point where a
// handler object would be created
Decision decide(const Request &r);

}  // namespace ckm_http
#endif
