/* kser_b200 -- the kser binary (kser.cc) over libckm.so: `kser_b200 [options] listen-port kmer-data-dir` */
#include "../../include/ckm_server.h"

int main(int argc, char **argv) { return ckm_kser_main(argc, argv); }
