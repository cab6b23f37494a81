#include "bam_writer.hpp"

#include <zlib.h>

#include <cstring>
#include <stdexcept>

namespace ab2 {

namespace {

inline void put_le16(std::string &o, uint32_t v) {
  o += static_cast<char>(v & 0xff);
  o += static_cast<char>((v >> 8) & 0xff);
}
inline void put_le32(std::string &o, uint32_t v) {
  put_le16(o, v & 0xffff);
  put_le16(o, v >> 16);
}

}  // namespace

void bgzf_append_block(const char *data, size_t n, int level, std::string &out) {
  if (n > kBgzfBlockPayload) throw std::runtime_error("bgzf block payload too large");
  // gzip member header with the BC extra subfield (SAM spec 4.1)
  static const unsigned char head[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
  const size_t at = out.size();
  out.append(reinterpret_cast<const char *>(head), 16);
  out.append(2, '\0');  // BSIZE, patched below
  z_stream zs;
  std::memset(&zs, 0, sizeof zs);
  if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK)
    throw std::runtime_error("deflateInit2 failed");
  const size_t bound = deflateBound(&zs, static_cast<uLong>(n));
  const size_t body = out.size();
  out.resize(body + bound);
  zs.next_in = reinterpret_cast<Bytef *>(const_cast<char *>(data));
  zs.avail_in = static_cast<uInt>(n);
  zs.next_out = reinterpret_cast<Bytef *>(&out[body]);
  zs.avail_out = static_cast<uInt>(bound);
  const int rc = deflate(&zs, Z_FINISH);
  const size_t clen = zs.total_out;
  deflateEnd(&zs);
  if (rc != Z_STREAM_END) throw std::runtime_error("deflate failed");
  out.resize(body + clen);
  const uint32_t crc = static_cast<uint32_t>(crc32(crc32(0L, Z_NULL, 0), reinterpret_cast<const Bytef *>(data),
                                                   static_cast<uInt>(n)));
  put_le32(out, crc);
  put_le32(out, static_cast<uint32_t>(n));
  const size_t total = out.size() - at;
  if (total > 0x10000) {
    // incompressible payload: store it instead (a stored deflate block costs 5 bytes)
    out.resize(body);
    out += static_cast<char>(1);  // BFINAL=1, BTYPE=00
    put_le16(out, static_cast<uint32_t>(n));
    put_le16(out, static_cast<uint32_t>(~n) & 0xffffu);
    out.append(data, n);
    put_le32(out, crc);
    put_le32(out, static_cast<uint32_t>(n));
  }
  const size_t bsize = out.size() - at - 1;
  out[at + 16] = static_cast<char>(bsize & 0xff);
  out[at + 17] = static_cast<char>((bsize >> 8) & 0xff);
}

void bgzf_append(const char *data, size_t n, int level, std::string &out) {
  while (n > 0) {
    const size_t k = n < kBgzfBlockPayload ? n : kBgzfBlockPayload;
    bgzf_append_block(data, k, level, out);
    data += k;
    n -= k;
  }
}

const std::string &bgzf_eof_marker() {
  static const std::string eof("\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00\x42\x43\x02\x00\x1b\x00\x03\x00"
                               "\x00\x00\x00\x00\x00\x00\x00\x00", 28);
  return eof;
}

std::string make_bam_header(const ChromLookup &cl, const std::string &text) {
  std::string h("BAM\1", 4);
  put_le32(h, static_cast<uint32_t>(text.size()));
  h += text;
  const size_t n_real = cl.names.size() >= 2 ? cl.names.size() - 2 : 0;
  put_le32(h, static_cast<uint32_t>(n_real));
  for (size_t i = 1; i + 1 < cl.names.size(); ++i) {
    put_le32(h, static_cast<uint32_t>(cl.names[i].size() + 1));
    h += cl.names[i];
    h += '\0';
    put_le32(h, cl.starts[i + 1] - cl.starts[i]);
  }
  return h;
}

void BgzfRecordPacker::add(const char *rec, size_t n) {
  if (block_.size() + n > kBgzfBlockPayload && !block_.empty()) {  // bgzf_flush_try
    bgzf_append_block(block_.data(), block_.size(), level_, out_);
    block_.clear();
  }
  while (n > kBgzfBlockPayload) {  // a record larger than a block is split (never the case for short reads)
    bgzf_append_block(rec, kBgzfBlockPayload, level_, out_);
    rec += kBgzfBlockPayload;
    n -= kBgzfBlockPayload;
  }
  block_.append(rec, n);
}

void BgzfRecordPacker::finish() {
  if (!block_.empty()) {
    bgzf_append_block(block_.data(), block_.size(), level_, out_);
    block_.clear();
  }
}

}  // namespace ab2
