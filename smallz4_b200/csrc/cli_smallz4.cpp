// cli_smallz4.cpp -- command-line front end with the reference's flags (reference: smallz4.cpp:166-326):
//   smallz4 [-0..-9] [-f] [-l] [-D dictionary] [-v] [-h] [input] [output]
// stdin/stdout when a name is missing or "-".  Compression runs on the GPU through include/smallz4.h.
#include "smallz4.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <unistd.h>

namespace
{
struct Io
{
  FILE* in; FILE* out;
  bool verbose;
  unsigned long long bytes_in, bytes_out;
  time_t start;
};

void fail(const char* msg, int code = 1)
{
  std::fprintf(stderr, "ERROR: %s\n", msg);
  std::exit(code);
}

size_t read_some(void* data, size_t want, void* user)
{
  Io* io = static_cast<Io*>(user);
  if (!data || want == 0) return 0;
  size_t got = std::fread(data, 1, want, io->in);
  io->bytes_in += got;
  return got;
}

void write_some(const void* data, size_t len, void* user)
{
  Io* io = static_cast<Io*>(user);
  if (!data || len == 0) return;
  std::fwrite(data, 1, len, io->out);
  io->bytes_out += len;
}

void usage(const char* prog)
{
  std::printf("smalLZ4 %s (B200): optimal-parse LZ4 compressor, output identical to smalLZ4 1.5\n\n"
              "Usage: %s [flags] [input] [output]\n"
              "  -0 ... -9   compression level, default 9 (-0 stores, -1..-%d greedy, -%d..-8 lazy/optimal with\n"
              "              short chains, -9 optimal parsing over all matches)\n"
              "  -f          overwrite an existing output file\n"
              "  -l          LZ4 legacy frame format\n"
              "  -D FILE     preload a dictionary (its last 64 KiB)\n"
              "  -v          verbose\n"
              "  -h          this text\n"
              "Reads STDIN / writes STDOUT when a file name is missing or \"-\".\n",
              smallz4::getVersion(), prog, smallz4::ShortChainsGreedy, smallz4::ShortChainsGreedy + 1);
}
}  // namespace

int main(int argc, char** argv)
{
  if (argc == 1 && isatty(fileno(stdin))) { usage(argv[0]); return 0; }

  unsigned short chain = 65535;      // smallz4.cpp:175
  bool overwrite = false, legacy = false;
  const char* dict_name = NULL;
  Io io = { stdin, stdout, false, 0, 0, 0 };

  int arg = 1;
  while (arg < argc && argv[arg][0] == '-' && argv[arg][1] != '\0')
  {
    bool takes_dict = false;
    for (const char* f = argv[arg] + 1; *f; ++f)
      switch (*f)
      {
        case 'h': usage(argv[0]); return 0;
        case 'f': overwrite = true; break;
        case 'l': legacy = true; break;
        case 'v': io.verbose = true; break;
        case 'D':
          if (arg + 1 >= argc) fail("no dictionary filename found");
          dict_name = argv[arg + 1];
          takes_dict = true;
          break;
        case '0': case '1': case '2': case '3': case '4': case '5': case '6': case '7': case '8':
          chain = (unsigned short)(*f - '0');           // smallz4.cpp:232-234
          break;
        case '9': chain = 65535; break;
        default: fail("unknown flag");
      }
    arg += takes_dict ? 2 : 1;
  }
  if (arg < argc && std::strcmp(argv[arg], "-") != 0)
  {
    io.in = std::fopen(argv[arg], "rb");
    if (!io.in) fail("file not found");
  }
  if (arg < argc) arg++;
  if (arg < argc && std::strcmp(argv[arg], "-") != 0)
  {
    if (!overwrite)
    {
      FILE* probe = std::fopen(argv[arg], "rb");
      if (probe) { std::fclose(probe); fail("output file already exists"); }
    }
    io.out = std::fopen(argv[arg], "wb");
    if (!io.out) fail("cannot create file");
  }
  if (legacy)                                            // smallz4.cpp:273-279
  {
    if (dict_name) fail("legacy format doesn't support dictionaries");
    if (chain == 0) fail("legacy format doesn't support uncompressed files");
  }

  std::vector<unsigned char> dict;
  if (dict_name)                                         // smallz4.cpp:283-304: the last 64 KiB
  {
    FILE* d = std::fopen(dict_name, "rb");
    if (!d) fail("cannot open dictionary");
    std::fseek(d, 0, SEEK_END);
    long size = std::ftell(d);
    long from = size < 65536 ? 0 : size - 65536;
    std::fseek(d, from, SEEK_SET);
    dict.resize((size_t)(size - from));
    if (!dict.empty() && std::fread(&dict[0], 1, dict.size(), d) != dict.size()) fail("cannot read dictionary");
    std::fclose(d);
  }

  io.start = std::time(NULL);
  try
  {
    smallz4::lz4(read_some, write_some, chain, dict, legacy, &io);
  }
  catch (const std::exception& e)
  {
    fail(e.what(), 2);
  }
  if (io.verbose && io.bytes_in > 0)
    std::fprintf(stderr, "%llu bytes => %llu bytes (%llu%%) after %d seconds\n", io.bytes_in, io.bytes_out,
                 100 * io.bytes_out / io.bytes_in, (int)(std::time(NULL) - io.start));
  if (io.out != stdout) std::fclose(io.out);
  return 0;
}
