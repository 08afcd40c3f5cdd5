// kq_oracle.cpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A row-at-a-time C++17 restatement of the physical operators of folkol/query-engines
// (kquerydiy/src/Main.kt, cited as Main.kt:N). Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library; the product
// (libkqgpu.so) never does and has no CPU fallback.
//
// PARITY UNPINNED BY THE REFERENCE: the reference ships no tests, no golden vectors and cannot be
// built or run in this image (no JVM, no build file; SURVEY.md §0 facts 1,4). The pins are
//  (1) the semantics rules R1-R12 read directly from Main.kt, followed line by
//      line below, (2) golden vectors derived by hand from the reference's only fixture
//      (kquerydiy/employee.csv) in tests/golden/, and (3) an independent numpy restatement in
//      tests/np_ref.py that must agree with this file.
// Operators the north star names but the reference lacks (filter, literal/binary expressions,
// SUM/MIN/COUNT) follow the documented extension rules E1-E8.
//
// Deliberately "boxed": every cell goes through Value, like ColumnVector.getValue(i): Any?.

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../include/kq_gen.h"

namespace {

// Status codes: identical numbering to include/kqgpu.h so tests compare them directly.
enum { OK = 0, E_ILLEGAL_STATE = 1, E_UNSUPPORTED = 2, E_ILLEGAL_ARGUMENT = 3, E_SQL = 4,
       E_NUMBER_FORMAT = 5, E_ARITHMETIC = 6 };
enum Type { T_F64 = 1, T_UTF8 = 2, T_I64 = 3, T_BOOL = 4, T_DATE32 = 5, T_I32 = 6 };
enum BinOp { EQ = 1, NE, LT, LE, GT, GE, AND, OR, ADD, SUB, MUL, DIV };
enum AggKind { A_MAX = 1, A_MIN = 2, A_SUM = 3, A_COUNT = 4 };

struct KqError : std::runtime_error {
    int code;
    KqError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

thread_local std::string g_last_error;

// ---- Value: the `Any?` of ColumnVector.getValue (Main.kt:24-27) ---------------------------------
struct Value {
    enum Kind : uint8_t { NUL = 0, F64, I64, BOOL, STR, DATE32 } kind = NUL;
    double d = 0;
    int64_t i = 0;
    std::string s;
    static Value null() { return Value(); }
    static Value f64(double v) { Value x; x.kind = F64; x.d = v; return x; }
    static Value i64(int64_t v) { Value x; x.kind = I64; x.i = v; return x; }
    static Value boolean(bool v) { Value x; x.kind = BOOL; x.i = v; return x; }
    static Value date32(int32_t v) { Value x; x.kind = DATE32; x.i = v; return x; }
    static Value str(const char* p, size_t n) { Value x; x.kind = STR; x.s.assign(p, n); return x; }
    bool isNull() const { return kind == NUL; }
};

// Java Double.equals/hashCode work on doubleToLongBits: every NaN is the same value and
// +0.0 != -0.0 (rule R7).
static inline uint64_t doubleToLongBits(double d) {
    if (d != d) return 0x7ff8000000000000ULL;
    uint64_t b; std::memcpy(&b, &d, 8); return b;
}

// ---- ColumnVector (Main.kt:24-27) and ArrowFieldVector (Main.kt:176-202) --------------------------
struct ColumnVector {
    virtual ~ColumnVector() {}
    virtual Value getValue(int64_t i) const = 0;
    virtual int64_t size() const = 0;
    virtual int type() const = 0;
};
using ColPtr = std::shared_ptr<ColumnVector>;

// Owns Arrow-layout buffers: validity bitmap LSB-first, int32 offsets (n+1), data bytes.
struct ArrowColumn : ColumnVector {
    int ty;
    int64_t n = 0;
    bool has_validity = false;
    std::vector<uint8_t> validity;
    std::vector<int32_t> offsets;
    std::vector<uint8_t> data;

    bool isNull(int64_t i) const {
        return has_validity && !((validity[(size_t)(i >> 3)] >> (i & 7)) & 1);
    }
    // Main.kt:178-197: the null check precedes the type dispatch (R1); Float8 -> Double,
    // VarChar -> String(bytes) (R2); anything else throws IllegalStateException (R3) — the
    // extension types of rule E1 are read here instead of throwing.
    Value getValue(int64_t i) const override {
        if (isNull(i)) return Value::null();
        switch (ty) {
            case T_F64: { double v; std::memcpy(&v, &data[(size_t)i * 8], 8); return Value::f64(v); }
            case T_I64: { int64_t v; std::memcpy(&v, &data[(size_t)i * 8], 8); return Value::i64(v); }
            case T_DATE32: { int32_t v; std::memcpy(&v, &data[(size_t)i * 4], 4); return Value::date32(v); }
            case T_I32: { int32_t v; std::memcpy(&v, &data[(size_t)i * 4], 4); return Value::i64(v); }
            case T_BOOL: return Value::boolean((data[(size_t)(i >> 3)] >> (i & 7)) & 1);
            case T_UTF8: {
                int32_t a = offsets[(size_t)i], b = offsets[(size_t)i + 1];
                return Value::str((const char*)data.data() + a, (size_t)(b - a));
            }
            default: throw KqError(E_ILLEGAL_STATE, "unreadable vector type");
        }
    }
    int64_t size() const override { return n; }   // field.valueCount, Main.kt:199-201
    int type() const override { return ty; }
};

// A row range of another column (used to partition a batch across threads like main()'s files).
struct SliceColumn : ColumnVector {
    ColPtr base; int64_t off, len;
    SliceColumn(ColPtr b, int64_t o, int64_t l) : base(std::move(b)), off(o), len(l) {}
    Value getValue(int64_t i) const override { return base->getValue(off + i); }
    int64_t size() const override { return len; }
    int type() const override { return base->type(); }
};

// ---- ArrowVectorBuilder / FieldVectorFactory (Main.kt:462-512) -------------------------------------
struct Builder {
    std::shared_ptr<ArrowColumn> col;
    int64_t utf8_pos = 0;
    explicit Builder(int ty, int64_t capacity) {
        // FieldVectorFactory.create (Main.kt:464-476) knows Float8 and VarChar only; E1 adds the rest.
        if (ty < T_F64 || ty > T_I32) throw KqError(E_ILLEGAL_STATE, "cannot create vector type");
        col = std::make_shared<ArrowColumn>();
        col->ty = ty;
        col->has_validity = true;
        col->validity.assign((size_t)((capacity + 7) / 8), 0);
        if (ty == T_UTF8) col->offsets.assign((size_t)capacity + 1, 0);
        else if (ty == T_BOOL) col->data.assign((size_t)((capacity + 7) / 8), 0);
        else col->data.assign((size_t)capacity * (ty == T_DATE32 || ty == T_I32 ? 4 : 8), 0);
    }
    // Main.kt:481-503. null -> setNull; Number -> toDouble() into Float8 (R12); toString() into Utf8.
    // Rows must be set in increasing order (they are, at every call site).
    void set(int64_t i, const Value& v) {
        ArrowColumn& c = *col;
        if (c.ty == T_UTF8) {
            if (!v.isNull()) {
                std::string s = toString(v);
                c.data.insert(c.data.end(), s.begin(), s.end());
                utf8_pos += (int64_t)s.size();
                c.validity[(size_t)(i >> 3)] |= (uint8_t)(1u << (i & 7));
            }
            c.offsets[(size_t)i + 1] = (int32_t)utf8_pos;
            return;
        }
        if (v.isNull()) return;  // validity bit stays 0
        c.validity[(size_t)(i >> 3)] |= (uint8_t)(1u << (i & 7));
        switch (c.ty) {
            case T_F64: {
                double d;
                if (v.kind == Value::F64) d = v.d;
                else if (v.kind == Value::I64 || v.kind == Value::DATE32) d = (double)v.i;  // Number.toDouble()
                else throw KqError(E_ILLEGAL_STATE, "Float8 builder: not a Number");        // Main.kt:497
                std::memcpy(&c.data[(size_t)i * 8], &d, 8); break;
            }
            case T_I64: {
                if (v.kind != Value::I64 && v.kind != Value::DATE32) throw KqError(E_ILLEGAL_STATE, "Int64 builder: not an integer");
                std::memcpy(&c.data[(size_t)i * 8], &v.i, 8); break;
            }
            case T_DATE32: case T_I32: {
                int32_t x = (int32_t)v.i; std::memcpy(&c.data[(size_t)i * 4], &x, 4); break;
            }
            case T_BOOL: if (v.i) c.data[(size_t)(i >> 3)] |= (uint8_t)(1u << (i & 7)); break;
            default: throw KqError(E_ILLEGAL_STATE, "builder type");                          // Main.kt:501
        }
    }
    static std::string toString(const Value& v) {
        switch (v.kind) {
            case Value::STR: return v.s;
            case Value::I64: case Value::DATE32: return std::to_string(v.i);
            case Value::BOOL: return v.i ? "true" : "false";
            case Value::F64: { char b[64]; snprintf(b, sizeof b, "%.17g", v.d); return b; }
            default: return "null";
        }
    }
    ColPtr build(int64_t valueCount) {   // setValueCount + build, Main.kt:505-511
        ArrowColumn& c = *col;
        c.n = valueCount;
        if (c.ty == T_UTF8) c.offsets.resize((size_t)valueCount + 1);
        return col;
    }
};

// ---- RecordBatch (Main.kt:56-61) -------------------------------------------------------------------
struct RecordBatch {
    std::vector<ColPtr> fields;
    int64_t explicit_rows = -1;
    int64_t rowCount() const {              // fields.first().size(), Main.kt:57
        if (fields.empty()) {
            if (explicit_rows < 0) throw KqError(E_ILLEGAL_STATE, "rowCount of empty batch");
            return explicit_rows;
        }
        return fields.front()->size();
    }
    const ColPtr& field(int i) const {
        if (i < 0 || i >= (int)fields.size()) throw KqError(E_ILLEGAL_STATE, "field index out of range");
        return fields[(size_t)i];
    }
};

// ---- Expression (Main.kt:448-450) -------------------------------------------------------------------
struct Expression {
    virtual ~Expression() {}
    virtual ColPtr evaluate(const RecordBatch& input) const = 0;
};
using ExprPtr = std::shared_ptr<Expression>;

// ColumnExpression (Main.kt:452-460): returns the input vector itself (alias, rule R4).
struct ColumnExpression : Expression {
    int i;
    explicit ColumnExpression(int idx) : i(idx) {}
    ColPtr evaluate(const RecordBatch& input) const override { return input.field(i); }
};

// A column that repeats one value (literal expressions are absent from the reference; E-rules).
struct LiteralColumn : ColumnVector {
    Value v; int64_t n; int ty;
    LiteralColumn(Value val, int64_t rows, int t) : v(std::move(val)), n(rows), ty(t) {}
    Value getValue(int64_t) const override { return v; }
    int64_t size() const override { return n; }
    int type() const override { return ty; }
};
struct LiteralExpression : Expression {
    Value v; int ty;
    LiteralExpression(Value val, int t) : v(std::move(val)), ty(t) {}
    ColPtr evaluate(const RecordBatch& input) const override {
        return std::make_shared<LiteralColumn>(v, input.rowCount(), ty);
    }
};

// Java's Double.parseDouble grammar (FloatingDecimal.readJavaFormatString), restated (rule R5):
// leading/trailing chars <= ' ' trimmed; optional sign; "NaN" | "Infinity" (case-sensitive);
// decimal digits with optional '.' and exponent; hex needs a 'p' exponent; one optional
// trailing d/D/f/F; "" and anything else throw NumberFormatException.
static double javaParseDouble(const std::string& in) {
    size_t b = 0, e = in.size();
    while (b < e && (unsigned char)in[b] <= ' ') b++;
    while (e > b && (unsigned char)in[e - 1] <= ' ') e--;
    std::string s = in.substr(b, e - b);
    auto fail = [&]() -> double { throw KqError(E_NUMBER_FORMAT, "For input string: \"" + in + "\""); };
    if (s.empty()) return fail();
    size_t p = 0; bool neg = false;
    if (s[p] == '+' || s[p] == '-') { neg = s[p] == '-'; p++; }
    std::string body = s.substr(p);
    if (body == "NaN") return std::nan("");
    if (body == "Infinity") return neg ? -INFINITY : INFINITY;
    if (body.empty()) return fail();
    // optional type suffix
    char last = body.back();
    if (last == 'd' || last == 'D' || last == 'f' || last == 'F') {
        bool hex = body.size() > 2 && body[0] == '0' && (body[1] == 'x' || body[1] == 'X');
        // in a hex literal 'd'/'f' are digits unless they follow the p-exponent
        if (!hex || body.find_first_of("pP") != std::string::npos) body.pop_back();
        if (body.empty()) return fail();
    }
    bool hex = body.size() > 2 && body[0] == '0' && (body[1] == 'x' || body[1] == 'X');
    size_t i = hex ? 2 : 0, nd = 0;
    auto isdig = [&](char c) { return hex ? std::isxdigit((unsigned char)c) != 0 : (c >= '0' && c <= '9'); };
    while (i < body.size() && isdig(body[i])) { i++; nd++; }
    if (i < body.size() && body[i] == '.') { i++; while (i < body.size() && isdig(body[i])) { i++; nd++; } }
    if (nd == 0) return fail();
    bool has_exp = false;
    if (i < body.size() && (hex ? (body[i] == 'p' || body[i] == 'P') : (body[i] == 'e' || body[i] == 'E'))) {
        has_exp = true; i++;
        if (i < body.size() && (body[i] == '+' || body[i] == '-')) i++;
        size_t ed = 0;
        while (i < body.size() && body[i] >= '0' && body[i] <= '9') { i++; ed++; }
        if (ed == 0) return fail();
    }
    if (i != body.size()) return fail();
    if (hex && !has_exp) return fail();
    double v = std::strtod(body.c_str(), nullptr);   // correctly rounded in glibc
    return neg ? -v : v;
}

// CastExpression (Main.kt:772-805): target Double only; null -> null; String.toDouble();
// a non-String source throws IllegalStateException (R5). Extension: Int64 -> Float64.
struct CastExpression : Expression {
    ExprPtr expr; int dataType;
    CastExpression(ExprPtr e, int t) : expr(std::move(e)), dataType(t) {}
    ColPtr evaluate(const RecordBatch& input) const override {
        ColPtr value = expr->evaluate(input);
        if (dataType != T_F64) throw KqError(E_ILLEGAL_STATE, "Cast to this type is not supported");  // Main.kt:799
        Builder builder(dataType, input.rowCount());                                               // Main.kt:780-781
        for (int64_t it = 0; it < value->size(); it++) {                                           // Main.kt:785
            Value vv = value->getValue(it);
            if (vv.isNull()) { builder.set(it, Value::null()); continue; }
            if (vv.kind == Value::STR) builder.set(it, Value::f64(javaParseDouble(vv.s)));         // Main.kt:791
            else if (vv.kind == Value::I64) builder.set(it, Value::f64((double)vv.i));             // extension
            else if (vv.kind == Value::F64) builder.set(it, vv);                                   // extension (identity)
            else throw KqError(E_ILLEGAL_STATE, "Cannot cast value to Double");                    // Main.kt:792
        }
        return builder.build(value->size());                                                        // Main.kt:802
    }
};

// BinaryExpression — ABSENT from the reference; rules E2 (same operand types), E3 (null in ->
// null out; AND/OR three-valued), E4 (separately rounded IEEE doubles, Int64 wraps, Int64 / 0 throws).
struct BinaryExpression : Expression {
    int op; ExprPtr l, r;
    BinaryExpression(int o, ExprPtr a, ExprPtr b) : op(o), l(std::move(a)), r(std::move(b)) {}

    static int cmp3(const Value& a, const Value& b, bool& unordered) {
        unordered = false;
        switch (a.kind) {
            case Value::F64:
                if (a.d != a.d || b.d != b.d) { unordered = true; return 0; }
                return a.d < b.d ? -1 : (a.d > b.d ? 1 : 0);
            case Value::I64: case Value::DATE32: case Value::BOOL:
                return a.i < b.i ? -1 : (a.i > b.i ? 1 : 0);
            case Value::STR: {   // unsigned byte order (= code point order for valid UTF-8)
                int c = std::memcmp(a.s.data(), b.s.data(), std::min(a.s.size(), b.s.size()));
                if (c != 0) return c < 0 ? -1 : 1;
                return a.s.size() < b.s.size() ? -1 : (a.s.size() > b.s.size() ? 1 : 0);
            }
            default: throw KqError(E_ILLEGAL_STATE, "compare");
        }
    }

    ColPtr evaluate(const RecordBatch& input) const override {
        ColPtr a = l->evaluate(input), b = r->evaluate(input);
        int ta = a->type(), tb = b->type();
        if (ta != tb) throw KqError(E_ILLEGAL_STATE, "binary operand types differ");     // E2
        int64_t n = input.rowCount();
        bool is_cmp = op >= EQ && op <= GE, is_logic = op == AND || op == OR;
        if (is_logic && ta != T_BOOL) throw KqError(E_ILLEGAL_STATE, "AND/OR need Bool operands");
        if (!is_cmp && !is_logic && ta != T_I64 && ta != T_F64)
            throw KqError(E_ILLEGAL_STATE, "math needs Int64 or Float64 operands");
        if (is_cmp && ta == T_I32) throw KqError(E_ILLEGAL_STATE, "compare on I32");
        int out_ty = (is_cmp || is_logic) ? T_BOOL : ta;
        Builder out(out_ty, n);
        for (int64_t i = 0; i < n; i++) {
            Value x = a->getValue(i), y = b->getValue(i);
            if (is_logic) {      // SQL three-valued logic (E3)
                bool xn = x.isNull(), yn = y.isNull();
                if (op == AND) {
                    if ((!xn && !x.i) || (!yn && !y.i)) out.set(i, Value::boolean(false));
                    else if (xn || yn) out.set(i, Value::null());
                    else out.set(i, Value::boolean(true));
                } else {
                    if ((!xn && x.i) || (!yn && y.i)) out.set(i, Value::boolean(true));
                    else if (xn || yn) out.set(i, Value::null());
                    else out.set(i, Value::boolean(false));
                }
                continue;
            }
            if (x.isNull() || y.isNull()) { out.set(i, Value::null()); continue; }       // E3
            if (is_cmp) {
                bool un; int c = cmp3(x, y, un);
                bool res;
                switch (op) {
                    case EQ: res = !un && c == 0; break;
                    case NE: res = un || c != 0; break;
                    case LT: res = !un && c < 0; break;
                    case LE: res = !un && c <= 0; break;
                    case GT: res = !un && c > 0; break;
                    default: res = !un && c >= 0; break;
                }
                out.set(i, Value::boolean(res));
                continue;
            }
            if (ta == T_F64) {
                volatile double p = x.d, q = y.d, res;        // volatile: no contraction into FMA (E4)
                switch (op) { case ADD: res = p + q; break; case SUB: res = p - q; break;
                              case MUL: res = p * q; break; default: res = p / q; break; }
                out.set(i, Value::f64(res));
            } else {
                uint64_t p = (uint64_t)x.i, q = (uint64_t)y.i; int64_t res;
                switch (op) {
                    case ADD: res = (int64_t)(p + q); break;   // Kotlin Long wraps
                    case SUB: res = (int64_t)(p - q); break;
                    case MUL: res = (int64_t)(p * q); break;
                    default:
                        if (y.i == 0) throw KqError(E_ARITHMETIC, "/ by zero");
                        res = (x.i == INT64_MIN && y.i == -1) ? INT64_MIN : x.i / y.i;  // JVM ldiv wraps
                        break;
                }
                out.set(i, Value::i64(res));
            }
        }
        return out.build(n);
    }
};

// ---- Accumulators (Main.kt:514-562) -----------------------------------------------------------------
struct Accumulator {
    virtual ~Accumulator() {}
    virtual void accumulate(const Value& v) = 0;
    virtual Value finalValue() const = 0;
};

// MaxAccumulator, Main.kt:538-562 (rule R9): ignore null; first non-null initialises; replace iff
// value > current (IEEE: comparisons with NaN are false). Byte|Double in the reference; Int64 and
// Date32 added by E1; anything else throws UnsupportedOperationException (Main.kt:548-550).
struct MaxAccumulator : Accumulator {
    Value value;
    void accumulate(const Value& v) override {
        if (v.isNull()) return;
        if (value.isNull()) { value = v; return; }
        bool isMax;
        switch (v.kind) {
            case Value::F64: isMax = v.d > value.d; break;
            case Value::I64: case Value::DATE32: isMax = v.i > value.i; break;
            default: throw KqError(E_UNSUPPORTED, "MAX is not implemented for this data type");
        }
        if (isMax) value = v;
    }
    Value finalValue() const override { return value; }
};
// MIN mirrors MAX with '<' (rule E5).
struct MinAccumulator : Accumulator {
    Value value;
    void accumulate(const Value& v) override {
        if (v.isNull()) return;
        if (value.isNull()) { value = v; return; }
        bool isMin;
        switch (v.kind) {
            case Value::F64: isMin = v.d < value.d; break;
            case Value::I64: case Value::DATE32: isMin = v.i < value.i; break;
            default: throw KqError(E_UNSUPPORTED, "MIN is not implemented for this data type");
        }
        if (isMin) value = v;
    }
    Value finalValue() const override { return value; }
};
// SUM skips nulls; all-null => null; Float64 sums in row order; Int64 wraps (rule E6).
struct SumAccumulator : Accumulator {
    Value value;
    void accumulate(const Value& v) override {
        if (v.isNull()) return;
        if (v.kind != Value::F64 && v.kind != Value::I64)
            throw KqError(E_UNSUPPORTED, "SUM is not implemented for this data type");
        if (value.isNull()) { value = v; return; }
        if (v.kind == Value::F64) { volatile double s = value.d + v.d; value.d = s; }
        else value.i = (int64_t)((uint64_t)value.i + (uint64_t)v.i);
    }
    Value finalValue() const override { return value; }
};
// COUNT(expr): non-null rows, Int64, never null (rule E7).
struct CountAccumulator : Accumulator {
    int64_t n = 0;
    void accumulate(const Value& v) override { if (!v.isNull()) n++; }
    Value finalValue() const override { return Value::i64(n); }
};

struct AggregateExpression {       // Main.kt:514-517
    int kind; ExprPtr expr;
    ExprPtr inputExpression() const { return expr; }
    std::unique_ptr<Accumulator> createAccumulator() const {
        switch (kind) {
            case A_MAX: return std::make_unique<MaxAccumulator>();
            case A_MIN: return std::make_unique<MinAccumulator>();
            case A_SUM: return std::make_unique<SumAccumulator>();
            case A_COUNT: return std::make_unique<CountAccumulator>();
            default: throw KqError(E_ILLEGAL_STATE, "Unsupported aggregate function");   // Main.kt:696
        }
    }
    int outputType(int input_type) const { return kind == A_COUNT ? T_I64 : input_type; }
};

// ---- operators ---------------------------------------------------------------------------------------
// ProjectionExec.execute for one batch, Main.kt:589-594 (rule R6).
static RecordBatch projectBatch(const std::vector<ExprPtr>& expr, const RecordBatch& batch) {
    RecordBatch out;
    out.explicit_rows = batch.rowCount();
    for (auto& e : expr) out.fields.push_back(e->evaluate(batch));
    return out;
}

static ColPtr takeRows(const ColPtr& c, const std::vector<int64_t>& sel) {
    Builder b(c->type(), (int64_t)sel.size());
    for (size_t k = 0; k < sel.size(); k++) b.set((int64_t)k, c->getValue(sel[k]));
    return b.build((int64_t)sel.size());
}

// FilterExec — ABSENT from the reference. Keeps rows whose predicate is TRUE, in input order;
// null predicate => dropped (rule E3).
static RecordBatch filterBatch(const ExprPtr& pred, const RecordBatch& batch, std::vector<int64_t>* sel_out) {
    ColPtr p = pred->evaluate(batch);
    if (p->type() != T_BOOL) throw KqError(E_ILLEGAL_STATE, "filter predicate is not Bool");
    std::vector<int64_t> sel;
    int64_t n = batch.rowCount();
    for (int64_t i = 0; i < n; i++) { Value v = p->getValue(i); if (!v.isNull() && v.i) sel.push_back(i); }
    RecordBatch out;
    out.explicit_rows = (int64_t)sel.size();
    for (auto& c : batch.fields) out.fields.push_back(takeRows(c, sel));
    if (sel_out) *sel_out = std::move(sel);
    return out;
}

// Group key = List<Any?> with Java boxed equality (Main.kt:621-627, rule R7).
struct RowKey {
    std::vector<Value> v;
    bool operator==(const RowKey& o) const {
        if (v.size() != o.v.size()) return false;
        for (size_t k = 0; k < v.size(); k++) {
            const Value &a = v[k], &b = o.v[k];
            if (a.kind != b.kind) return false;
            switch (a.kind) {
                case Value::NUL: break;
                case Value::F64: if (doubleToLongBits(a.d) != doubleToLongBits(b.d)) return false; break;
                case Value::STR: if (a.s != b.s) return false; break;
                default: if (a.i != b.i) return false; break;
            }
        }
        return true;
    }
};
struct RowKeyHash {
    size_t operator()(const RowKey& k) const {   // List.hashCode shape: 31*h + elem.hashCode
        uint64_t h = 1;
        for (const Value& a : k.v) {
            uint64_t e;
            switch (a.kind) {
                case Value::NUL: e = 0; break;
                case Value::F64: e = doubleToLongBits(a.d); e ^= e >> 32; break;
                case Value::STR: e = std::hash<std::string>()(a.s); break;
                default: e = (uint64_t)a.i ^ ((uint64_t)a.i >> 32); break;
            }
            h = 31 * h + e;
        }
        return (size_t)kq_mix64(h);
    }
};

// HashAggregateExec (Main.kt:605-660). update() is one iteration of the drain loop (617-634);
// finalize() is the single-batch emit (635-650). Output row order here is unordered_map iteration
// order — like the reference's HashMap order it is an artefact, never asserted (rules R10, E8).
struct HashAggregateExec {
    ExprPtr pred;                              // optional fused FilterExec below the aggregate
    std::vector<ExprPtr> groupExpr;
    std::vector<AggregateExpression> aggregateExpr;
    std::unordered_map<RowKey, std::vector<std::unique_ptr<Accumulator>>, RowKeyHash> map;
    std::vector<int> groupTypes, aggInputTypes;

    void update(const RecordBatch& in) {
        RecordBatch filtered;
        const RecordBatch* batch = &in;
        if (pred) { filtered = filterBatch(pred, in, nullptr); batch = &filtered; }
        std::vector<ColPtr> groupKeys, aggrInputValues;
        for (auto& g : groupExpr) groupKeys.push_back(g->evaluate(*batch));                              // Main.kt:618
        for (auto& a : aggregateExpr) aggrInputValues.push_back(a.inputExpression()->evaluate(*batch));  // Main.kt:619
        if (groupTypes.empty() && aggInputTypes.empty()) {
            for (auto& c : groupKeys) groupTypes.push_back(c->type());
            for (auto& c : aggrInputValues) aggInputTypes.push_back(c->type());
        }
        int64_t n = batch->rowCount();
        for (int64_t rowIndex = 0; rowIndex < n; rowIndex++) {                                           // Main.kt:620
            RowKey rowKey;
            for (auto& c : groupKeys) rowKey.v.push_back(c->getValue(rowIndex));                         // Main.kt:621-626
            auto it = map.find(rowKey);                                                                   // getOrPut, Main.kt:627
            if (it == map.end()) {
                std::vector<std::unique_ptr<Accumulator>> accs;
                for (auto& a : aggregateExpr) accs.push_back(a.createAccumulator());
                it = map.emplace(std::move(rowKey), std::move(accs)).first;
            }
            for (size_t k = 0; k < it->second.size(); k++)                                               // Main.kt:628-631
                it->second[k]->accumulate(aggrInputValues[k]->getValue(rowIndex));                       // nulls included (R8)
        }
    }

    RecordBatch finalize() {
        if (groupTypes.empty() && aggInputTypes.empty() && (!groupExpr.empty() || !aggregateExpr.empty()) && map.empty()) {
            // No batch was ever seen: the output schema comes from the planner in the reference;
            // here the caller must have fed at least one (possibly empty) batch to fix types.
            throw KqError(E_ILLEGAL_STATE, "finalize before any update: output types unknown");
        }
        int64_t rows = (int64_t)map.size();                                                               // Main.kt:637
        std::vector<Builder> builders;
        for (size_t k = 0; k < groupExpr.size(); k++) builders.emplace_back(groupTypes[k], rows);
        for (size_t k = 0; k < aggregateExpr.size(); k++)
            builders.emplace_back(aggregateExpr[k].outputType(aggInputTypes[k]), rows);
        int64_t rowIndex = 0;
        for (auto& entry : map) {                                                                         // Main.kt:639-647
            for (size_t k = 0; k < groupExpr.size(); k++) builders[k].set(rowIndex, entry.first.v[k]);
            for (size_t k = 0; k < aggregateExpr.size(); k++)
                builders[groupExpr.size() + k].set(rowIndex, entry.second[k]->finalValue());
            rowIndex++;
        }
        RecordBatch out;
        out.explicit_rows = rows;
        for (auto& b : builders) out.fields.push_back(b.build(rows));
        return out;                                                                                       // one batch, Main.kt:649-650
    }
};

// ---- synthetic tables (include/kq_gen.h) --------------------------------------------------------------
static ColPtr generateColumn(const kq_gen_spec& sp, uint64_t seed, int64_t r0, int64_t r1) {
    int64_t n = r1 - r0;
    auto c = std::make_shared<ArrowColumn>();
    c->n = n;
    switch (sp.kind) {
        case KQ_GEN_I64_UNIFORM: c->ty = T_I64; c->data.resize((size_t)n * 8); break;
        case KQ_GEN_F64_UNIFORM: case KQ_GEN_F64_INT: case KQ_GEN_F64_STEP: c->ty = T_F64; c->data.resize((size_t)n * 8); break;
        case KQ_GEN_UTF8_DICT: c->ty = T_UTF8; c->offsets.resize((size_t)n + 1); c->data.resize((size_t)n * sp.dict_width); break;
        case KQ_GEN_DATE32_UNIFORM: c->ty = T_DATE32; c->data.resize((size_t)n * 4); break;
        case KQ_GEN_BOOL: c->ty = T_BOOL; c->data.assign((size_t)((n + 7) / 8), 0); break;
        default: throw KqError(E_ILLEGAL_ARGUMENT, "unknown generator kind");
    }
    if (sp.null_per_10k > 0) { c->has_validity = true; c->validity.assign((size_t)((n + 7) / 8), 0); }
    for (int64_t i = 0; i < n; i++) {
        int64_t row = r0 + i;
        uint64_t h = kq_gen_hash(seed, sp.col_id, row);
        if (c->has_validity && !kq_gen_is_null(seed, sp.col_id, row, sp.null_per_10k))
            c->validity[(size_t)(i >> 3)] |= (uint8_t)(1u << (i & 7));
        switch (sp.kind) {
            case KQ_GEN_I64_UNIFORM: { int64_t v = kq_gen_i64(h, sp.ilo, sp.ihi); std::memcpy(&c->data[(size_t)i * 8], &v, 8); break; }
            case KQ_GEN_F64_UNIFORM: { double v = kq_gen_f64_uniform(h, sp.flo, sp.fhi); std::memcpy(&c->data[(size_t)i * 8], &v, 8); break; }
            case KQ_GEN_F64_INT: { double v = (double)kq_gen_i64(h, sp.ilo, sp.ihi); std::memcpy(&c->data[(size_t)i * 8], &v, 8); break; }
            case KQ_GEN_F64_STEP: { double v = kq_gen_f64_step(h, sp.ilo, sp.ihi, sp.fhi); std::memcpy(&c->data[(size_t)i * 8], &v, 8); break; }
            case KQ_GEN_UTF8_DICT: {
                int64_t code = (int64_t)(h % (uint64_t)sp.dict_count);
                std::memcpy(&c->data[(size_t)i * sp.dict_width], sp.dict + code * sp.dict_width, (size_t)sp.dict_width);
                c->offsets[(size_t)i] = (int32_t)(i * sp.dict_width);
                c->offsets[(size_t)i + 1] = (int32_t)((i + 1) * sp.dict_width);
                break;
            }
            case KQ_GEN_DATE32_UNIFORM: { int32_t v = (int32_t)kq_gen_i64(h, sp.ilo, sp.ihi); std::memcpy(&c->data[(size_t)i * 4], &v, 4); break; }
            case KQ_GEN_BOOL: if ((int64_t)(h % 10000ULL) < sp.ilo) c->data[(size_t)(i >> 3)] |= (uint8_t)(1u << (i & 7)); break;
        }
    }
    if (sp.kind == KQ_GEN_UTF8_DICT && n == 0) c->offsets[0] = 0;
    return c;
}

// Materialise any ColumnVector into Arrow buffers (for handing results back over the C API).
static std::shared_ptr<ArrowColumn> materialize(const ColPtr& c) {
    if (auto a = std::dynamic_pointer_cast<ArrowColumn>(c)) return a;
    Builder b(c->type(), c->size());
    for (int64_t i = 0; i < c->size(); i++) b.set(i, c->getValue(i));
    return std::static_pointer_cast<ArrowColumn>(b.build(c->size()));
}

static RecordBatch sliceBatch(const RecordBatch& b, int64_t off, int64_t len) {
    RecordBatch s; s.explicit_rows = len;
    for (auto& c : b.fields) s.fields.push_back(std::make_shared<SliceColumn>(c, off, len));
    return s;
}

}  // namespace

// ======================================================================================================
// C API (ctypes). Same shapes as include/kqgpu.h with a ko_ prefix, so tests drive both identically.
// ======================================================================================================
struct ko_col { ColPtr c; };
struct ko_batch { RecordBatch b; };
struct ko_expr { ExprPtr e; };
// ---- CsvDataSource.scan + ReaderIterator.createBatch (Main.kt:251-273, 276-357) -----------------------------
// The reference hands tokenising to univocity-parsers (com.univocity:univocity-parsers, NOT under /root/reference, no
// pinned version: PARITY UNPINNED) with delimiter/line-separator detection, skipEmptyLines and header extraction
// (Main.kt:289-296, 322), then stores getValue(name, "").trim() of every field (Main.kt:262-264). Restated as rules
// C1-C9 (query-engines_b200/csrc/kq_csv.cu header), here as a plain character-at-a-time tokenizer:
//   C1 separator "\n" (CRLF: the CR goes with the terminator) or lone "\r" when the text has no "\n"; delimiter = most
//      frequent of , ; TAB | outside quotes in the first record (found with all four counting as delimiters).
//   C2 a '"' opens a quoted section only as the first non-blank byte of a field; inside, "" is a literal quote and a single
//      '"' closes; every other '"' is data. What follows the closing quote of a value is dropped.
//   C3 lines without a byte are skipped.  C4 header record names the columns.  C5 values are trimmed (<= 0x20).
//   C6 all columns Utf8, never null; missing field = "".  C7 projection by file column index.  C8 surplus fields ignored.
struct CsvText {
    std::vector<std::vector<std::string>> records;
    char delim = ',', term = '\n';
};
static std::string csv_trim(const std::string& v) {
    size_t a = 0, b = v.size();
    while (a < b && (unsigned char)v[a] <= 0x20) a++;
    while (b > a && (unsigned char)v[b - 1] <= 0x20) b--;
    return v.substr(a, b - a);
}
static std::string csv_field_value(const std::string& raw) {
    std::string t = csv_trim(raw);
    if (t.empty() || t[0] != '"') return t;
    std::string v;
    for (size_t i = 1; i < t.size(); i++) {
        if (t[i] == '"') {
            if (i + 1 < t.size() && t[i + 1] == '"') { v += '"'; i++; }
            else break;                       // closing quote: the rest of the field is dropped
        } else v += t[i];
    }
    return csv_trim(v);                       // String.trim() on the parsed value (Main.kt:263)
}
// Rule C2 one character at a time. Where a character stands: at the start of a field (only blanks so far), in an unquoted
// value, inside a quoted section, or right behind a '"' met inside a quoted section (it closed the section unless the next
// character is another '"'). `delims` are the bytes that separate fields (one byte when the delimiter is known; all four
// candidates while it is being detected).
struct CsvQuoteRule {
    enum Where { FieldStart, Unquoted, Quoted, QuoteInQuoted } at = FieldStart;
    std::string delims;
    char term;
    CsvQuoteRule(const std::string& d, char t) : delims(d), term(t) {}
    bool separates(char c) const { return c == term || delims.find(c) != std::string::npos; }
    // consumes c; true when c stands outside quotes (so a delimiter or terminator there is a real one)
    bool outside(char c) {
        switch (at) {
            case FieldStart:
                if (separates(c)) return true;
                if (c == '"') { at = Quoted; return false; }
                if ((unsigned char)c > 0x20) at = Unquoted;
                return true;
            case Unquoted:
                if (separates(c)) at = FieldStart;
                return true;                                       // a '"' here is data
            case Quoted:
                if (c == '"') at = QuoteInQuoted;
                return false;
            default:                                               // QuoteInQuoted
                if (c == '"') { at = Quoted; return false; }       // "" = a literal quote
                at = separates(c) ? FieldStart : Unquoted;         // the quote before c closed the section
                return true;
        }
    }
};
// lines: split at terminators outside quotes
// (first_record_only: stop behind the first line that is not empty — all that delimiter detection looks at)
static std::vector<std::string> csv_lines(const std::string& all, const std::string& delims, char term, bool whole_text, bool first_record_only = false) {
    std::vector<std::string> lines;
    std::string cur;
    CsvQuoteRule q(delims, term);
    for (char c : all) {
        if (q.outside(c) && c == term) {
            const bool blank = cur.empty() || (term == '\n' && cur == "\r");
            lines.push_back(cur); cur.clear();
            if (first_record_only && !blank) return lines;
        }
        else cur += c;
    }
    if (whole_text && q.at == CsvQuoteRule::Quoted) throw KqError(E_ILLEGAL_STATE, "CSV text ends inside a quoted field");
    if (!cur.empty()) lines.push_back(cur);
    return lines;
}
static CsvText csv_tokenize(const uint8_t* text, int64_t n) {
    CsvText out;
    std::string all((const char*)text, (size_t)n);
    if (all.find('\n') == std::string::npos && all.find('\r') != std::string::npos) out.term = '\r';
    auto empty_line = [&](std::string& line) {
        if (out.term == '\n' && !line.empty() && line.back() == '\r') line.pop_back();     // CRLF
        return line.empty();                                                               // skipEmptyLines (Main.kt:293)
    };
    // delimiter detection on the first record (Main.kt:291), found with every candidate counting as a delimiter
    const std::string cand = ",;\t|";
    for (std::string& line : csv_lines(all, cand, out.term, false, true)) {
        if (empty_line(line)) continue;
        long cnt[4] = {0, 0, 0, 0};
        CsvQuoteRule q(cand, out.term);
        for (char c : line) if (q.outside(c)) for (int k = 0; k < 4; k++) cnt[k] += c == cand[(size_t)k];
        int best = 0;
        for (int k = 1; k < 4; k++) if (cnt[k] > cnt[best]) best = k;
        out.delim = cand[(size_t)best];
        break;
    }
    const std::string delim(1, out.delim);
    for (std::string& line : csv_lines(all, delim, out.term, true)) {
        if (empty_line(line)) continue;
        std::vector<std::string> fields;
        std::string cur;
        CsvQuoteRule q(delim, out.term);
        for (char c : line) {
            if (q.outside(c) && c == out.delim) { fields.push_back(csv_field_value(cur)); cur.clear(); }
            else cur += c;
        }
        fields.push_back(csv_field_value(cur));
        out.records.push_back(std::move(fields));
    }
    return out;
}

struct ko_hashagg { HashAggregateExec h; };

#define KO_TRY(body) \
    try { body; return OK; } \
    catch (const KqError& e) { g_last_error = e.what(); return e.code; } \
    catch (const std::exception& e) { g_last_error = e.what(); return E_ILLEGAL_STATE; }

extern "C" {

const char* ko_last_error() { return g_last_error.c_str(); }

int ko_column_new(int type, int64_t n, const uint8_t* validity, const int32_t* offsets,
                  const void* data, int64_t data_bytes, ko_col** out) {
    KO_TRY({
        if (type < T_F64 || type > T_I32) throw KqError(E_ILLEGAL_STATE, "unknown column type");
        auto c = std::make_shared<ArrowColumn>();
        c->ty = type; c->n = n;
        if (validity) { c->has_validity = true; c->validity.assign(validity, validity + (n + 7) / 8); }
        size_t bytes;
        if (type == T_UTF8) {
            c->offsets.assign(offsets, offsets + n + 1);
            bytes = (size_t)data_bytes;
        } else if (type == T_BOOL) bytes = (size_t)((n + 7) / 8);
        else bytes = (size_t)n * ((type == T_DATE32 || type == T_I32) ? 4 : 8);
        if (bytes) c->data.assign((const uint8_t*)data, (const uint8_t*)data + bytes);
        *out = new ko_col{c};
    })
}
int ko_column_sizes(ko_col* col, int64_t* n, int64_t* data_bytes, int64_t* null_count) {
    KO_TRY({
        auto a = materialize(col->c); col->c = a;
        if (n) *n = a->n;
        if (data_bytes) *data_bytes = a->ty == T_UTF8 ? (a->n ? a->offsets[(size_t)a->n] - a->offsets[0] : 0) : (int64_t)a->data.size();
        if (null_count) { int64_t k = 0; for (int64_t i = 0; i < a->n; i++) k += a->isNull(i); *null_count = k; }
    })
}
int ko_column_type(ko_col* col) { return col->c->type(); }
int ko_column_download(ko_col* col, uint8_t* validity, int32_t* offsets, void* data) {
    KO_TRY({
        auto a = materialize(col->c); col->c = a;
        if (validity) for (int64_t i = 0; i < (a->n + 7) / 8; i++) validity[i] = 0;
        if (validity) for (int64_t i = 0; i < a->n; i++) if (!a->isNull(i)) validity[i >> 3] |= (uint8_t)(1u << (i & 7));
        if (a->ty == T_UTF8) {
            int32_t base = a->n ? a->offsets[0] : 0;
            if (offsets) for (int64_t i = 0; i <= a->n; i++) offsets[i] = a->offsets[(size_t)i] - base;
            if (data && a->n) std::memcpy(data, a->data.data() + base, (size_t)(a->offsets[(size_t)a->n] - base));
        } else if (data && !a->data.empty()) std::memcpy(data, a->data.data(), a->data.size());
    })
}
int ko_column_free(ko_col* c) { delete c; return OK; }

// CsvDataSource.inferSchema (Main.kt:328-356): names separated by '\n'
int ko_csv_header(const uint8_t* text, int64_t nbytes, int has_headers, char* names, size_t cap, int* ncols, char* delimiter) {
    KO_TRY({
        CsvText t = csv_tokenize(text, nbytes);
        std::string all;
        const size_t k = t.records.empty() ? 0 : t.records[0].size();
        for (size_t i = 0; i < k; i++) all += (has_headers ? t.records[0][i] : "field_" + std::to_string(i + 1)) + "\n";
        if (all.size() + 1 > cap) throw KqError(E_ILLEGAL_ARGUMENT, "names buffer too small");
        std::memcpy(names, all.c_str(), all.size() + 1);
        *ncols = (int)k;
        if (delimiter) *delimiter = t.delim;
    })
}
// CsvDataSource.scan (Main.kt:304-326) + createBatch (Main.kt:251-273), all records in one batch (rule C9)
int ko_csv_scan(const uint8_t* text, int64_t nbytes, int has_headers, const int* projection, int nproj, ko_batch** out) {
    KO_TRY({
        CsvText t = csv_tokenize(text, nbytes);
        const int file_cols = t.records.empty() ? 0 : (int)t.records[0].size();
        std::vector<int> proj;
        if (nproj) proj.assign(projection, projection + nproj);
        else for (int i = 0; i < file_cols; i++) proj.push_back(i);
        for (int c : proj) if (c < 0 || c >= file_cols) throw KqError(E_ILLEGAL_ARGUMENT, "projected CSV column out of range");
        const size_t skip = has_headers && !t.records.empty() ? 1 : 0;
        const int64_t rows = (int64_t)(t.records.size() - skip);
        auto b = new ko_batch();
        b->b.explicit_rows = rows;
        for (int c : proj) {
            auto col = std::make_shared<ArrowColumn>();
            col->ty = T_UTF8; col->n = rows;
            col->offsets.push_back(0);
            for (size_t r = skip; r < t.records.size(); r++) {
                const std::vector<std::string>& rec = t.records[r];
                const std::string v = (size_t)c < rec.size() ? rec[(size_t)c] : std::string();     // getValue(name, "") (Main.kt:263)
                col->data.insert(col->data.end(), v.begin(), v.end());
                col->offsets.push_back((int32_t)col->data.size());
            }
            b->b.fields.push_back(col);
        }
        *out = b;
    })
}

int ko_batch_create(ko_col* const* cols, int ncols, int64_t n_rows, ko_batch** out) {
    KO_TRY({
        auto b = new ko_batch();
        b->b.explicit_rows = n_rows;
        for (int i = 0; i < ncols; i++) b->b.fields.push_back(cols[i]->c);
        for (int i = 1; i < ncols; i++)
            if (cols[i]->c->size() != cols[0]->c->size()) { delete b; throw KqError(E_ILLEGAL_ARGUMENT, "column lengths differ"); }
        *out = b;
    })
}
int ko_batch_num_rows(ko_batch* b, int64_t* n) { KO_TRY({ *n = b->b.rowCount(); }) }
int ko_batch_num_columns(ko_batch* b) { return (int)b->b.fields.size(); }
int ko_batch_column(ko_batch* b, int i, ko_col** out) { KO_TRY({ *out = new ko_col{b->b.field(i)}; }) }
int ko_batch_free(ko_batch* b) { delete b; return OK; }

ko_expr* ko_expr_column(int i) { return new ko_expr{std::make_shared<ColumnExpression>(i)}; }
ko_expr* ko_expr_literal_f64(double v) { return new ko_expr{std::make_shared<LiteralExpression>(Value::f64(v), T_F64)}; }
ko_expr* ko_expr_literal_i64(int64_t v) { return new ko_expr{std::make_shared<LiteralExpression>(Value::i64(v), T_I64)}; }
ko_expr* ko_expr_literal_bool(int v) { return new ko_expr{std::make_shared<LiteralExpression>(Value::boolean(v != 0), T_BOOL)}; }
ko_expr* ko_expr_literal_date32(int32_t v) { return new ko_expr{std::make_shared<LiteralExpression>(Value::date32(v), T_DATE32)}; }
ko_expr* ko_expr_literal_utf8(const char* p, int32_t len) { return new ko_expr{std::make_shared<LiteralExpression>(Value::str(p, (size_t)len), T_UTF8)}; }
ko_expr* ko_expr_literal_null(int type) { return new ko_expr{std::make_shared<LiteralExpression>(Value::null(), type)}; }
ko_expr* ko_expr_binary(int op, ko_expr* l, ko_expr* r) { return new ko_expr{std::make_shared<BinaryExpression>(op, l->e, r->e)}; }
ko_expr* ko_expr_cast(ko_expr* e, int type) { return new ko_expr{std::make_shared<CastExpression>(e->e, type)}; }
void ko_expr_free(ko_expr* e) { delete e; }

int ko_expr_evaluate(ko_expr* e, ko_batch* in, ko_col** out) { KO_TRY({ *out = new ko_col{e->e->evaluate(in->b)}; }) }

int ko_project(ko_expr* const* exprs, int n, ko_batch* in, ko_batch** out) {
    KO_TRY({
        std::vector<ExprPtr> ex; for (int i = 0; i < n; i++) ex.push_back(exprs[i]->e);
        auto b = new ko_batch(); b->b = projectBatch(ex, in->b); *out = b;
    })
}
int ko_filter(ko_expr* pred, ko_batch* in, ko_batch** out, ko_col** selection) {
    KO_TRY({
        std::vector<int64_t> sel;
        auto b = new ko_batch(); b->b = filterBatch(pred->e, in->b, &sel); *out = b;
        if (selection) {
            Builder sb(T_I32, (int64_t)sel.size());
            for (size_t k = 0; k < sel.size(); k++) sb.set((int64_t)k, Value::i64(sel[k]));
            auto sc = std::static_pointer_cast<ArrowColumn>(sb.build((int64_t)sel.size()));
            sc->has_validity = false;
            *selection = new ko_col{sc};
        }
    })
}
int ko_filter_project(ko_expr* pred, ko_expr* const* exprs, int n, ko_batch* in, ko_batch** out) {
    KO_TRY({
        std::vector<ExprPtr> ex; for (int i = 0; i < n; i++) ex.push_back(exprs[i]->e);
        RecordBatch f = filterBatch(pred->e, in->b, nullptr);     // FilterExec, then ProjectionExec
        auto b = new ko_batch(); b->b = projectBatch(ex, f); *out = b;
    })
}

int ko_hashagg_create(ko_expr* pred, ko_expr* const* group_exprs, int ngroup, const int* agg_kinds,
                      ko_expr* const* agg_inputs, int nagg, ko_hashagg** out) {
    KO_TRY({
        auto h = new ko_hashagg();
        if (pred) h->h.pred = pred->e;
        for (int i = 0; i < ngroup; i++) h->h.groupExpr.push_back(group_exprs[i]->e);
        for (int i = 0; i < nagg; i++) {
            if (agg_kinds[i] < A_MAX || agg_kinds[i] > A_COUNT) { delete h; throw KqError(E_ILLEGAL_STATE, "Unsupported aggregate function"); }
            h->h.aggregateExpr.push_back(AggregateExpression{agg_kinds[i], agg_inputs[i]->e});
        }
        *out = h;
    })
}
int ko_hashagg_update(ko_hashagg* h, ko_batch* in) { KO_TRY({ h->h.update(in->b); }) }
int ko_hashagg_finalize(ko_hashagg* h, ko_batch** out) { KO_TRY({ auto b = new ko_batch(); b->b = h->h.finalize(); *out = b; }) }
int ko_hashagg_free(ko_hashagg* h) { delete h; return OK; }

int ko_generate(const kq_gen_spec* specs, int ncols, uint64_t seed, int64_t row_begin, int64_t row_end, ko_batch** out) {
    KO_TRY({
        auto b = new ko_batch(); b->b.explicit_rows = row_end - row_begin;
        for (int i = 0; i < ncols; i++) b->b.fields.push_back(generateColumn(specs[i], seed, row_begin, row_end));
        *out = b;
    })
}

// ---- partition -> partial -> merge on `nthreads` host threads, the shape of main() (Main.kt:1309-1325).
// Used only for the reported CPU baseline. Returns the number of output rows in *out_rows.
int ko_filter_project_mt(ko_expr* pred, ko_expr* const* exprs, int n, ko_batch* in, int nthreads, int64_t* out_rows) {
    KO_TRY({
        std::vector<ExprPtr> ex; for (int i = 0; i < n; i++) ex.push_back(exprs[i]->e);
        int64_t rows = in->b.rowCount();
        std::vector<std::thread> th; std::vector<int64_t> cnt((size_t)nthreads, 0);
        std::atomic<int> failed{0};
        for (int t = 0; t < nthreads; t++) th.emplace_back([&, t]() {
            try {
                int64_t a = rows * t / nthreads, b = rows * (t + 1) / nthreads;
                RecordBatch s = sliceBatch(in->b, a, b - a);
                RecordBatch f = pred ? filterBatch(pred->e, s, nullptr) : s;
                RecordBatch p = projectBatch(ex, f);
                for (auto& c : p.fields) materialize(c);
                cnt[(size_t)t] = p.rowCount();
            } catch (...) { failed = 1; }
        });
        for (auto& x : th) x.join();
        if (failed) throw KqError(E_ILLEGAL_STATE, "worker failed");
        int64_t total = 0; for (auto c : cnt) total += c;
        *out_rows = total;
    })
}

int ko_hashagg_mt(ko_expr* pred, ko_expr* const* group_exprs, int ngroup, const int* agg_kinds,
                  ko_expr* const* agg_inputs, int nagg, ko_batch* in, int nthreads, ko_batch** out) {
    KO_TRY({
        int64_t rows = in->b.rowCount();
        std::vector<RecordBatch> partials((size_t)nthreads);
        std::vector<std::thread> th; std::atomic<int> failed{0};
        for (int t = 0; t < nthreads; t++) th.emplace_back([&, t]() {
            try {
                HashAggregateExec h;
                if (pred) h.pred = pred->e;
                for (int i = 0; i < ngroup; i++) h.groupExpr.push_back(group_exprs[i]->e);
                for (int i = 0; i < nagg; i++) h.aggregateExpr.push_back(AggregateExpression{agg_kinds[i], agg_inputs[i]->e});
                int64_t a = rows * t / nthreads, b = rows * (t + 1) / nthreads;
                h.update(sliceBatch(in->b, a, b - a));
                partials[(size_t)t] = h.finalize();
            } catch (...) { failed = 1; }
        });
        for (auto& x : th) x.join();
        if (failed) throw KqError(E_ILLEGAL_STATE, "worker failed");
        // merge query over the concatenated partials: MAX(max), MIN(min), SUM(sum), SUM(count) (Main.kt:1320)
        HashAggregateExec m;
        for (int i = 0; i < ngroup; i++) m.groupExpr.push_back(std::make_shared<ColumnExpression>(i));
        for (int i = 0; i < nagg; i++) {
            int k = agg_kinds[i] == A_COUNT ? A_SUM : agg_kinds[i];
            m.aggregateExpr.push_back(AggregateExpression{k, std::make_shared<ColumnExpression>(ngroup + i)});
        }
        for (auto& p : partials) m.update(p);
        auto b = new ko_batch(); b->b = m.finalize(); *out = b;
    })
}

double ko_parse_double(const char* s, int32_t len, int* status) {
    try { *status = OK; return javaParseDouble(std::string(s, (size_t)len)); }
    catch (const KqError& e) { g_last_error = e.what(); *status = e.code; return 0; }
}

}  // extern "C"
