// wire.cpp -- flat little-endian file format for keys and ciphertexts.
//
// The reference has no serialisation at all (no serde, no file I/O: keys and ciphertexts only ever live in RAM;
// SURVEY section 5), so the format is ours: the C-ABI layouts of include/tfhe_b200.h (u32 words, little endian) behind a
// 64-byte header.  It doubles as the checkpoint / resume format for keys and as the exchange format between the
// oracle and the device in cross-process parity runs.
//
//   offset  size  field
//        0     8  magic "TFHEB200"
//        8     4  version (1)
//       12     4  kind (TFHE_B200_FILE_*)
//       16    24  n, N, l, bgbit, ks_t, ks_basebit   (6 x u32: the parameter set the payload belongs to)
//       40     8  count   (records: ciphertexts / TRGSW samples; 1 for keys)
//       48     8  payload bytes
//       56     8  FNV-1a 64 of the payload
//       64     -  payload
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>
#include <string>
#include "../../include/tfhe_b200.h"

namespace {
constexpr uint32_t n = 635, N = 1024, L = 3, BGBIT = 6, KS_T = 8, KS_BB = 2;
const char MAGIC[8] = {'T', 'F', 'H', 'E', 'B', '2', '0', '0'};
thread_local std::string g_err;

uint64_t record_bytes(int kind) {
    switch (kind) {
    case TFHE_B200_FILE_SECRET: return n + N;                                  // s0 bytes, then s1 bytes (0/1)
    case TFHE_B200_FILE_BK: return (uint64_t)n * 2 * L * 2 * N * 4;
    case TFHE_B200_FILE_KSK: return (uint64_t)N * KS_T * 3 * (n + 1) * 4;
    case TFHE_B200_FILE_TLWE0: return (uint64_t)(n + 1) * 4;
    case TFHE_B200_FILE_TLWE1: return (uint64_t)(N + 1) * 4;
    case TFHE_B200_FILE_TRLWE: return (uint64_t)2 * N * 4;
    case TFHE_B200_FILE_TRGSW: return (uint64_t)2 * L * 2 * N * 4;
    default: return 0;
    }
}
uint64_t fnv1a(const uint8_t* p, uint64_t len) {
    uint64_t h = 0xCBF29CE484222325ull;
    for (uint64_t i = 0; i < len; i++) { h ^= p[i]; h *= 0x100000001B3ull; }
    return h;
}
void put32(uint8_t* p, uint32_t v) { for (int i = 0; i < 4; i++) p[i] = (uint8_t)(v >> (8 * i)); }
void put64(uint8_t* p, uint64_t v) { for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (8 * i)); }
uint32_t get32(const uint8_t* p) { uint32_t v = 0; for (int i = 0; i < 4; i++) v |= (uint32_t)p[i] << (8 * i); return v; }
uint64_t get64(const uint8_t* p) { uint64_t v = 0; for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i); return v; }
int fail(int code, const std::string& m) { g_err = m; return code; }

struct Header { int kind; uint64_t count, bytes, hash; };
int read_header(FILE* f, const char* path, Header* h) {
    uint8_t b[64];
    if (fread(b, 1, 64, f) != 64) return fail(TFHE_B200_ERR_IO, std::string(path) + ": short header");
    if (memcmp(b, MAGIC, 8) != 0) return fail(TFHE_B200_ERR_IO, std::string(path) + ": bad magic");
    if (get32(b + 8) != 1) return fail(TFHE_B200_ERR_IO, std::string(path) + ": unsupported version");
    h->kind = (int)get32(b + 12);
    const uint32_t want[6] = {n, N, L, BGBIT, KS_T, KS_BB};
    for (int i = 0; i < 6; i++)
        if (get32(b + 16 + 4 * i) != want[i]) return fail(TFHE_B200_ERR_IO, std::string(path) + ": parameter set mismatch");
    h->count = get64(b + 40); h->bytes = get64(b + 48); h->hash = get64(b + 56);
    const uint64_t rb = record_bytes(h->kind);
    if (rb == 0 || h->bytes != rb * h->count) return fail(TFHE_B200_ERR_IO, std::string(path) + ": inconsistent kind / count / length");
    return TFHE_B200_OK;
}
}  // namespace

extern "C" {

const char* tfhe_b200_file_last_error(void) { return g_err.c_str(); }

// payload: `count` records of the kind's flat layout.  On a little-endian host (every CUDA host) the u32 layouts are
// written as they lie in memory.
int tfhe_b200_file_write(const char* path, int kind, const void* payload, uint64_t count) {
    const uint64_t rb = record_bytes(kind);
    if (!path || rb == 0 || (!payload && count)) return fail(TFHE_B200_ERR_PARAM, "file_write: bad argument");
    if ((kind == TFHE_B200_FILE_SECRET || kind == TFHE_B200_FILE_BK || kind == TFHE_B200_FILE_KSK) && count != 1)
        return fail(TFHE_B200_ERR_PARAM, "file_write: key files hold exactly one record");
    const uint64_t bytes = rb * count;
    uint8_t h[64];
    memset(h, 0, sizeof h);
    memcpy(h, MAGIC, 8);
    put32(h + 8, 1); put32(h + 12, (uint32_t)kind);
    const uint32_t prm[6] = {n, N, L, BGBIT, KS_T, KS_BB};
    for (int i = 0; i < 6; i++) put32(h + 16 + 4 * i, prm[i]);
    put64(h + 40, count); put64(h + 48, bytes); put64(h + 56, fnv1a((const uint8_t*)payload, bytes));
    const std::string tmp = std::string(path) + ".tmp";
    // secret keys: created with mode 0600 and O_EXCL (never through an existing, possibly world-readable, file); others 0644
    remove(tmp.c_str());
    const int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_CLOEXEC, kind == TFHE_B200_FILE_SECRET ? 0600 : 0644);
    FILE* f = fd >= 0 ? fdopen(fd, "wb") : nullptr;
    if (!f) { if (fd >= 0) close(fd); return fail(TFHE_B200_ERR_IO, tmp + ": cannot open for writing"); }
    bool ok = fwrite(h, 1, 64, f) == 64 && (bytes == 0 || fwrite(payload, 1, bytes, f) == bytes);
    ok = (fclose(f) == 0) && ok;
    if (!ok) { remove(tmp.c_str()); return fail(TFHE_B200_ERR_IO, tmp + ": write failed"); }
    if (rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); return fail(TFHE_B200_ERR_IO, std::string(path) + ": rename failed"); }
    return TFHE_B200_OK;
}
int tfhe_b200_file_info(const char* path, int* kind, uint64_t* count, uint64_t* payload_bytes) {
    if (!path) return fail(TFHE_B200_ERR_PARAM, "file_info: bad argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(TFHE_B200_ERR_IO, std::string(path) + ": cannot open");
    Header h;
    const int rc = read_header(f, path, &h);
    fclose(f);
    if (rc) return rc;
    if (kind) *kind = h.kind;
    if (count) *count = h.count;
    if (payload_bytes) *payload_bytes = h.bytes;
    return TFHE_B200_OK;
}
// reads the payload into a caller buffer of exactly payload_bytes (from file_info); verifies kind and checksum
int tfhe_b200_file_read(const char* path, int kind, void* payload, uint64_t payload_bytes) {
    if (!path || (!payload && payload_bytes)) return fail(TFHE_B200_ERR_PARAM, "file_read: bad argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(TFHE_B200_ERR_IO, std::string(path) + ": cannot open");
    Header h;
    int rc = read_header(f, path, &h);
    if (!rc && h.kind != kind) rc = fail(TFHE_B200_ERR_IO, std::string(path) + ": holds a different kind of object");
    if (!rc && h.bytes != payload_bytes) rc = fail(TFHE_B200_ERR_PARAM, "file_read: buffer size does not match the file");
    if (!rc && h.bytes && fread(payload, 1, h.bytes, f) != h.bytes) rc = fail(TFHE_B200_ERR_IO, std::string(path) + ": truncated payload");
    if (!rc && fgetc(f) != EOF) rc = fail(TFHE_B200_ERR_IO, std::string(path) + ": trailing bytes");
    fclose(f);
    if (!rc && fnv1a((const uint8_t*)payload, h.bytes) != h.hash) rc = fail(TFHE_B200_ERR_IO, std::string(path) + ": checksum mismatch");
    return rc;
}
}
