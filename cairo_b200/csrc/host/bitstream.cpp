// bitstream.cpp -- the bit FIFO of the public API (semantics of bitstream.cpp:8-405 and
// memory.cpp:40-78 of the reference: LSB-first inside each byte, capacity in bits, reads that
// overwrite exactly `count` bits of the destination and leave its other bits alone).
#include <string.h>

#include <new>

#include "evx1.h"

namespace evx {

namespace {

// Copies n bits from (src, bit offset so) to (dst, bit offset d0); every other bit of dst is preserved.
void copy_bits(const uint8 *src, uint32 so, uint32 n, uint8 *dst, uint32 d0)
{
    if (!n) return;
    if (((so | d0) & 7) == 0)
    {
        uint32 whole = n >> 3;
        if (whole) memcpy(dst + (d0 >> 3), src + (so >> 3), whole);
        so += whole << 3; d0 += whole << 3; n &= 7;
    }
    while (n)
    {
        uint32 sb = so & 7, db = d0 & 7;
        uint32 room = 8 - (sb > db ? sb : db);
        uint32 take = n < room ? n : room;
        uint32 mask = (1u << take) - 1u;
        uint8 bits = (uint8) ((src[so >> 3] >> sb) & mask);
        uint8 *t = dst + (d0 >> 3);
        *t = (uint8) ((*t & ~(mask << db)) | (bits << db));
        so += take; d0 += take; n -= take;
    }
}

inline uint32 ceil_bytes(uint32 bits) { return (bits + 7u) >> 3; }

}  // namespace

bit_stream::bit_stream() : read_index(0), write_index(0), data_capacity(0), data_store(0) {}

bit_stream::bit_stream(uint32 size) : read_index(0), write_index(0), data_capacity(0), data_store(0) { resize_capacity(size); }

bit_stream::bit_stream(void *bytes, uint32 size) : read_index(0), write_index(0), data_capacity(0), data_store(0) { assign(bytes, size); }

bit_stream::~bit_stream() { clear(); }

uint8 *bit_stream::query_data() const { return data_store; }
uint32 bit_stream::query_capacity() const { return data_capacity << 3; }
uint32 bit_stream::query_occupancy() const { return write_index - read_index; }
uint32 bit_stream::query_byte_occupancy() const { return ceil_bytes(query_occupancy()); }

uint32 bit_stream::resize_capacity(uint32 size_in_bits)
{
    clear();
    if (!size_in_bits) return 0;
    uint32 bytes = ceil_bytes(size_in_bits);
    data_store = new (std::nothrow) uint8[bytes];
    if (!data_store) return 0;
    memset(data_store, 0, bytes);            // the reference leaves this uninitialised
    data_capacity = bytes;
    return size_in_bits;
}

evx_status bit_stream::assign(void *bytes, uint32 size)
{
    if (!bytes || !size) return EVX_ERROR_INVALIDARG;
    clear();
    data_store = new (std::nothrow) uint8[size];
    if (!data_store) return EVX_ERROR_OUTOFMEMORY;
    memcpy(data_store, bytes, size);
    read_index = 0;
    write_index = size << 3;
    data_capacity = size;
    return EVX_SUCCESS;
}

void bit_stream::seek(uint32 offset)
{
    if (read_index + offset >= write_index) read_index = write_index;
    read_index += offset;
}

void bit_stream::clear()
{
    empty();
    delete[] data_store;
    data_store = 0;
    data_capacity = 0;
}

void bit_stream::empty() { write_index = 0; read_index = 0; }
bool bit_stream::is_empty() const { return write_index == read_index; }
bool bit_stream::is_full() const { return write_index == query_capacity(); }

evx_status bit_stream::write_bit(uint8 value)
{
    if (write_index + 1 > query_capacity()) return EVX_ERROR_CAPACITY_LIMIT;
    uint8 *d = data_store + (write_index >> 3);
    uint32 k = write_index & 7;
    *d = (uint8) ((*d & ~(1u << k)) | ((value & 1u) << k));
    write_index++;
    return EVX_SUCCESS;
}

evx_status bit_stream::write_byte(uint8 value)
{
    if (write_index + 8 > query_capacity()) return EVX_ERROR_CAPACITY_LIMIT;
    copy_bits(&value, 0, 8, data_store, write_index);
    write_index += 8;
    return EVX_SUCCESS;
}

evx_status bit_stream::write_bits(void *data, uint32 bit_count)
{
    if (!data || !bit_count) return EVX_ERROR_INVALIDARG;
    if (write_index + bit_count > query_capacity()) return EVX_ERROR_CAPACITY_LIMIT;
    copy_bits(static_cast<const uint8 *>(data), 0, bit_count, data_store, write_index);
    write_index += bit_count;
    return EVX_SUCCESS;
}

evx_status bit_stream::write_bytes(void *data, uint32 byte_count) { return write_bits(data, byte_count << 3); }

evx_status bit_stream::peek_bit(void *data)
{
    if (!data) return EVX_ERROR_INVALIDARG;
    if (read_index >= write_index) return EVX_ERROR_INVALID_RESOURCE;
    uint8 *d = static_cast<uint8 *>(data);
    *d = (uint8) ((*d & 0xFE) | ((data_store[read_index >> 3] >> (read_index & 7)) & 1u));
    return EVX_SUCCESS;
}

evx_status bit_stream::peek_byte(void *data)
{
    if (!data) return EVX_ERROR_INVALIDARG;
    if (read_index + 8 > write_index) return EVX_ERROR_INVALID_RESOURCE;
    copy_bits(data_store, read_index, 8, static_cast<uint8 *>(data), 0);
    return EVX_SUCCESS;
}

evx_status bit_stream::peek_bits(void *data, uint32 count)
{
    if (!data || !count) return EVX_ERROR_INVALIDARG;
    if (read_index + count > write_index) return EVX_ERROR_INVALID_RESOURCE;
    copy_bits(data_store, read_index, count, static_cast<uint8 *>(data), 0);
    return EVX_SUCCESS;
}

evx_status bit_stream::peek_bytes(void *data, uint32 count) { return peek_bits(data, count << 3); }

evx_status bit_stream::read_bit(void *data)
{
    evx_status r = peek_bit(data);
    if (r == EVX_SUCCESS) read_index++;
    return r;
}

evx_status bit_stream::read_byte(void *data)
{
    evx_status r = peek_byte(data);
    if (r == EVX_SUCCESS) read_index += 8;
    return r;
}

evx_status bit_stream::read_bits(void *data, uint32 count)
{
    evx_status r = peek_bits(data, count);
    if (r == EVX_SUCCESS) read_index += count;
    return r;
}

evx_status bit_stream::read_bytes(void *data, uint32 count) { return read_bits(data, count << 3); }

}  // namespace evx
