// series.cu — native ingest of the reference's input files `data/ChargerXX_all_{train,eval,test}_fix.csv` (host code only).
//
// The environment reads 8 of the 21 columns (shems_LU1.jl:251-260, 268-279) and the reference re-parses the whole file with
// CSV.read on every reset and every step (:217, :265).  Here the file is parsed once, by column NAME (header order of
// Data_preparation_v2.ipynb cell 35), straight into the [8][nrows] float32 layout shems_create takes; Float64 text values are
// rounded to Float32 exactly as `env.state.x = df[idx, :col]` rounds them.  Bool columns (true/false) and `missing` fields
// (empty) only occur in columns the environment ignores.
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "common.h"

static const char* const kCols[SHEMS_SERIES_COLS] = {"soc_ev", "h_countdown", "electkwh", "PV_generation", "p_buy", "hour_cos", "hour_sin", "season"};

static std::string trim(const std::string& x) {
  size_t a = 0, b = x.size();
  while (a < b && (x[a] == ' ' || x[a] == '\t' || x[a] == '\r' || x[a] == '"' || (unsigned char)x[a] == 0xEF || (unsigned char)x[a] == 0xBB ||
                   (unsigned char)x[a] == 0xBF)) ++a;  // also strips a UTF-8 byte-order mark in front of the first header name
  while (b > a && (x[b - 1] == ' ' || x[b - 1] == '\t' || x[b - 1] == '\r' || x[b - 1] == '\n' || x[b - 1] == '"')) --b;
  return x.substr(a, b - a);
}
static void split(const std::string& line, std::vector<std::string>& out) {
  out.clear();
  size_t start = 0;
  for (;;) {
    const size_t c = line.find(',', start);
    if (c == std::string::npos) { out.push_back(trim(line.substr(start))); break; }
    out.push_back(trim(line.substr(start, c - start)));
    start = c + 1;
  }
}

// series_out == NULL: only count the rows (call again with a buffer of [8][*nrows_out] floats, capacity = its row count)
extern "C" int32_t shems_series_from_csv(const char* path, float* series_out, int32_t capacity, int32_t* nrows_out) {
  REQUIRE(path && nrows_out, SHEMS_ERR_INVALID, "shems_series_from_csv: NULL argument");
  FILE* f = fopen(path, "rb");
  REQUIRE(f, SHEMS_ERR_INVALID, "shems_series_from_csv: cannot open %s (%s)", path, strerror(errno));
  std::string text;
  char buf[1 << 16];
  size_t got;
  while ((got = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, got);
  fclose(f);
  std::vector<std::string> fields;
  size_t pos = 0;
  auto next_line = [&](std::string& line) -> bool {
    if (pos >= text.size()) return false;
    const size_t e = text.find('\n', pos);
    line = text.substr(pos, e == std::string::npos ? std::string::npos : e - pos);
    pos = e == std::string::npos ? text.size() : e + 1;
    return true;
  };
  std::string line;
  REQUIRE(next_line(line), SHEMS_ERR_INVALID, "shems_series_from_csv: %s is empty", path);
  split(line, fields);
  int col[SHEMS_SERIES_COLS];
  for (int k = 0; k < SHEMS_SERIES_COLS; ++k) {
    col[k] = -1;
    for (size_t j = 0; j < fields.size(); ++j) if (fields[j] == kCols[k]) { col[k] = (int)j; break; }
    REQUIRE(col[k] >= 0, SHEMS_ERR_KEY, "shems_series_from_csv: column :%s not found in %s (ArgumentError of df[idx, :%s], shems_LU1.jl:251-260)", kCols[k], path, kCols[k]);
  }
  int32_t n = 0;
  long lineno = 1;
  while (next_line(line)) {
    ++lineno;
    if (trim(line).empty()) continue;
    split(line, fields);
    if (series_out) {
      REQUIRE(n < capacity, SHEMS_ERR_INVALID, "shems_series_from_csv: %s has more than %d data rows", path, capacity);
      for (int k = 0; k < SHEMS_SERIES_COLS; ++k) {
        REQUIRE((size_t)col[k] < fields.size() && !fields[col[k]].empty() && fields[col[k]] != "missing", SHEMS_ERR_INVALID,
                "shems_series_from_csv: %s line %ld: column :%s is missing", path, lineno, kCols[k]);
        char* end = nullptr;
        const double v = strtod(fields[col[k]].c_str(), &end);
        REQUIRE(end && *end == '\0', SHEMS_ERR_INVALID, "shems_series_from_csv: %s line %ld: cannot parse '%s' in column :%s", path, lineno,
                fields[col[k]].c_str(), kCols[k]);
        series_out[(size_t)k * capacity + n] = (float)v;  // Float64 -> Float32, round to nearest even like Julia's convert
      }
    }
    ++n;
  }
  REQUIRE(n >= 1, SHEMS_ERR_INVALID, "shems_series_from_csv: %s has no data rows", path);
  *nrows_out = n;
  return SHEMS_OK;
}
