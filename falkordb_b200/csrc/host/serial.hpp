// serial.hpp -- C++ mirror of the reference's RDB encode / decode of the GraphBLAS objects on the path
// (graph/src/graph/graphblas/serialization.rs:13-90 Writer / Reader; matrix.rs:428-546; vector.rs:150-420), written against
// the serialization entry points of the C ABI exactly as the Rust files are (GxB_Container_*, GxB_unload_Matrix_into_Container /
// GxB_load_Matrix_from_Container, GxB_Vector_unload / load, GxB_Vector_serialize / deserialize, GrB_Type_get_String,
// GxB_Type_from_name).  `Stream` stands for both traits: the ordered items Redis' RDB I/O would carry.  Decode errors are
// std::runtime_error with the reference's messages (its Err(String)).
#pragma once
#include "matrix.hpp"
#include <cstdlib>
#include <cstring>
#include <deque>

namespace fdb {

class Stream {
  public:
    enum Kind { Unsigned, Signed, Buffer };
    struct Item { Kind kind; uint64_t u; int64_t s; std::vector<uint8_t> b; };
    std::deque<Item> items;

    void write_unsigned(uint64_t v) { items.push_back(Item{Unsigned, v, 0, {}}); }
    void write_signed(int64_t v) { items.push_back(Item{Signed, 0, v, {}}); }
    void write_buffer(const void *p, size_t n) {
        Item it{Buffer, 0, 0, {}};
        if (n) it.b.assign((const uint8_t *)p, (const uint8_t *)p + n);
        items.push_back(std::move(it));
    }
    uint64_t read_unsigned() { Item it = take(Unsigned, "unsigned"); return it.u; }
    int64_t read_signed() { Item it = take(Signed, "signed"); return it.s; }
    std::vector<uint8_t> read_buffer() { Item it = take(Buffer, "buffer"); return std::move(it.b); }
    bool empty() const { return items.empty(); }

  private:
    Item take(Kind k, const char *what) {
        if (items.empty()) throw std::runtime_error(std::string("unexpected end of stream reading ") + what);
        Item it = std::move(items.front());
        items.pop_front();
        if (it.kind != k) throw std::runtime_error(std::string("stream item is not a ") + what);
        return it;
    }
};

// <Vector<bool> as Encode<19>>::encode, vector.rs:241-309: unload to an array, write (array, type name + NUL, n, bytes, handling),
// load it back so the vector stays usable
inline void encode_payload_vector(GrB_Vector v, Stream &w) {
    void *arr = nullptr;
    GrB_Type type = nullptr;
    uint64_t n_entries = 0, n_bytes = 0;
    int handling = 0;
    grb_ok(GxB_Vector_unload(v, &arr, &type, &n_entries, &n_bytes, &handling, nullptr), "GxB_Vector_unload");
    char t_name[GxB_MAX_NAME_LEN] = {0};
    grb_ok(GrB_Type_get_String(type, t_name, GrB_NAME), "GrB_Type_get_String");
    size_t t_len = strnlen(t_name, GxB_MAX_NAME_LEN) + 1;
    if (t_len > GxB_MAX_NAME_LEN) t_len = GxB_MAX_NAME_LEN;
    w.write_buffer(arr, (size_t)n_bytes);
    w.write_buffer(t_name, t_len);
    w.write_unsigned(n_entries);
    w.write_unsigned(n_bytes);
    w.write_signed(handling);
    grb_ok(GxB_Vector_load(v, &arr, type, n_entries, n_bytes, handling, nullptr), "GxB_Vector_load");
}

// <Vector<bool> as Decode<19>>::decode, vector.rs:311-413, with its validation of the untrusted payload
inline GrB_Vector decode_payload_vector(Stream &r) {
    std::vector<uint8_t> arr_data = r.read_buffer();
    std::vector<uint8_t> type_name = r.read_buffer();
    uint64_t n_entries = r.read_unsigned();
    uint64_t n_bytes = r.read_unsigned();
    int handling = (int)r.read_signed();
    if (n_bytes != arr_data.size())
        throw std::runtime_error("Vector decode: declared byte length " + std::to_string(n_bytes) + " does not match buffer length " +
                                 std::to_string(arr_data.size()));
    if (type_name.empty() || type_name.back() != 0 || std::memchr(type_name.data(), 0, type_name.size() - 1))
        throw std::runtime_error("Vector decode: type name is not NUL-terminated");
    GrB_Type type = nullptr;
    GrB_Info info = GxB_Type_from_name(&type, (const char *)type_name.data());
    if (info != GrB_SUCCESS) throw std::runtime_error("Vector decode: GxB_Type_from_name failed: " + std::to_string((int)info));
    GrB_Vector v = nullptr;
    grb_ok(GrB_Vector_new(&v, type, 0), "GrB_Vector_new");
    void *arr = nullptr;
    if (n_bytes) {
        arr = std::malloc((size_t)n_bytes);                 // GxB_Vector_load takes ownership; freed with GxB_init's free (the process allocator here)
        if (!arr) { GrB_Vector_free(&v); throw std::bad_alloc(); }
        std::memcpy(arr, arr_data.data(), (size_t)n_bytes);
    }
    info = GxB_Vector_load(v, &arr, type, n_entries, n_bytes, handling, nullptr);
    if (info != GrB_SUCCESS) {
        if (arr) std::free(arr);
        GrB_Vector_free(&v);
        throw std::runtime_error("Vector decode: GxB_Vector_load failed: " + std::to_string((int)info));
    }
    return v;
}

static const size_t CONTAINER_STRUCT_SIZE = sizeof(struct GxB_Container_struct);    // 608, mod.rs:14191

// <Matrix<T> as Encode<19>>::encode, matrix.rs:508-546
template <class T>
inline void encode_matrix(const Matrix<T> &m, Stream &w) {
    GxB_Container c = nullptr;
    grb_ok(GxB_Container_new(&c), "GxB_Container_new");
    try {
        grb_ok(GxB_unload_Matrix_into_Container(m.inner(), c, nullptr), "GxB_unload_Matrix_into_Container");
        w.write_buffer(c, CONTAINER_STRUCT_SIZE);
        encode_payload_vector(c->x, w);
        encode_payload_vector(c->h, w);
        encode_payload_vector(c->p, w);
        encode_payload_vector(c->i, w);
        encode_payload_vector(c->b, w);
        grb_ok(GxB_load_Matrix_from_Container(m.inner(), c, nullptr), "GxB_load_Matrix_from_Container");
    } catch (...) {
        GxB_Container_free(&c);
        throw;
    }
    GxB_Container_free(&c);
}

// <Matrix<T> as Decode<19>>::decode, matrix.rs:428-506
template <class T>
inline Matrix<T> decode_matrix(Stream &r) {
    std::vector<uint8_t> bytes = r.read_buffer();
    if (bytes.size() < CONTAINER_STRUCT_SIZE)
        throw std::runtime_error("container buffer too small: " + std::to_string(bytes.size()) + " bytes < " +
                                 std::to_string(CONTAINER_STRUCT_SIZE) + " bytes required");
    GxB_Container c = nullptr;
    grb_ok(GxB_Container_new(&c), "GxB_Container_new");
    GrB_Matrix m = nullptr;
    try {
        // the empty vectors Container_new made are released before the struct bytes overwrite their handles
        GrB_Vector_free(&c->x); GrB_Vector_free(&c->h); GrB_Vector_free(&c->p); GrB_Vector_free(&c->i); GrB_Vector_free(&c->b);
        std::memcpy((void *)c, bytes.data(), CONTAINER_STRUCT_SIZE);
        c->x = nullptr; c->h = nullptr; c->b = nullptr; c->i = nullptr; c->p = nullptr; c->Y = nullptr;
        c->x = decode_payload_vector(r);
        c->h = decode_payload_vector(r);
        c->p = decode_payload_vector(r);
        c->i = decode_payload_vector(r);
        c->b = decode_payload_vector(r);
        grb_ok(GrB_Matrix_new(&m, GrB_BOOL, 0, 0), "GrB_Matrix_new");
        GrB_Matrix_set_INT32(m, GxB_SPARSE | GxB_HYPERSPARSE, GxB_SPARSITY_CONTROL);
        GrB_Matrix_set_INT32(m, GrB_ROWMAJOR, GrB_STORAGE_ORIENTATION_HINT);
        grb_ok(GxB_load_Matrix_from_Container(m, c, nullptr), "GxB_load_Matrix_from_Container");
        grb_ok(GrB_Matrix_wait(m, GrB_MATERIALIZE), "GrB_Matrix_wait");
    } catch (...) {
        if (m) GrB_Matrix_free(&m);
        GxB_Container_free(&c);
        throw;
    }
    GxB_Container_free(&c);
    return Matrix<T>::adopt(m, false);
}

// Vector::encode_blob / decode_blob, vector.rs:157-195: the GxB_Vector_serialize form C's RDB tensor section uses for a multi-edge
// pair's id list -- a BOOL vector of size GrB_INDEX_MAX whose INDICES are the edge ids
inline void encode_id_blob(const std::vector<uint64_t> &ids, uint64_t size, Stream &w) {
    GrB_Vector v = nullptr;
    grb_ok(GrB_Vector_new(&v, GrB_BOOL, size), "GrB_Vector_new");
    void *blob = nullptr;
    GrB_Index blob_size = 0;
    try {
        for (uint64_t id : ids) grb_ok(GrB_Vector_setElement_BOOL(v, true, id), "GrB_Vector_setElement_BOOL");
        grb_ok(GxB_Vector_serialize(&blob, &blob_size, v, nullptr), "GxB_Vector_serialize");
        w.write_buffer(blob, (size_t)blob_size);
    } catch (...) {
        if (blob) std::free(blob);
        GrB_Vector_free(&v);
        throw;
    }
    std::free(blob);
    GrB_Vector_free(&v);
}
inline std::vector<uint64_t> decode_id_blob(Stream &r) {
    std::vector<uint8_t> blob = r.read_buffer();
    GrB_Vector v = nullptr;
    grb_ok(GxB_Vector_deserialize(&v, nullptr, blob.data(), blob.size(), nullptr), "GxB_Vector_deserialize");
    std::vector<uint64_t> ids;
    GxB_Iterator it = nullptr;
    try {                                                       // Vector::iter, vector.rs:546-600
        grb_ok(GxB_Iterator_new(&it), "GxB_Iterator_new");
        grb_ok(GxB_Vector_Iterator_attach(it, v, nullptr), "GxB_Vector_Iterator_attach");
        GrB_Info info = GxB_Vector_Iterator_seek(it, 0);
        while (info == GrB_SUCCESS) {
            ids.push_back(GxB_Vector_Iterator_getIndex(it));
            info = GxB_Vector_Iterator_next(it);
        }
    } catch (...) {
        if (it) GxB_Iterator_free(&it);
        GrB_Vector_free(&v);
        throw;
    }
    GxB_Iterator_free(&it);
    GrB_Vector_free(&v);
    return ids;
}

} // namespace fdb
