// TEST-ONLY host build of the decimal / fixed-point headers (phanotate_b200/csrc/*.cuh are
// __host__ __device__).  Lets tests/test_dec_host.py compare every arithmetic primitive with
// Python's decimal (libmpdec) on the CPU box.  Not part of the product, never loaded by it.
#include "../../phanotate_b200/csrc/fxpow.cuh"
#include "../../phanotate_b200/csrc/frepr.cuh"
#include "../../phanotate_b200/csrc/hold.cuh"
#include "../../phanotate_b200/csrc/dec2double.cuh"

struct TDec {
    u64 lo, hi;
    i32 e, neg;
};
static Dec in(const TDec& t) {
    Dec d;
    d.c.w[0] = (u32)t.lo;
    d.c.w[1] = (u32)(t.lo >> 32);
    d.c.w[2] = (u32)t.hi;
    d.c.w[3] = (u32)(t.hi >> 32);
    d.e = t.e;
    d.neg = t.neg;
    return d;
}
static TDec out(const Dec& d) {
    TDec t;
    t.lo = d.c.w[0] | ((u64)d.c.w[1] << 32);
    t.hi = d.c.w[2] | ((u64)d.c.w[3] << 32);
    t.e = d.e;
    t.neg = d.neg;
    return t;
}
extern "C" {
void t_binop(int op, int n, const TDec* a, const TDec* b, int prec, TDec* o) {
    for (int i = 0; i < n; i++) {
        Dec x = in(a[i]), y = in(b[i]), r;
        switch (op) {
            case 0: r = dec_add(x, y, prec); break;
            case 1: r = dec_sub(x, y, prec); break;
            case 2: r = dec_mul(x, y, prec); break;
            default: r = dec_div(x, y, prec); break;
        }
        o[i] = out(r);
    }
}
void t_powi(int n, const TDec* a, const u32* nn, int prec, TDec* o) {
    for (int i = 0; i < n; i++) o[i] = out(dec_powi(in(a[i]), nn[i], prec));
}
// y given as Decimal
void t_powr(int n, const TDec* a, const TDec* y, int prec, TDec* o, int* okv) {
    for (int i = 0; i < n; i++) {
        bool ok1, ok2;
        Dec yy = in(y[i]);
        int yneg = yy.neg;
        yy.neg = 0;
        Fx Y = fx_from_dec(yy, &ok1);
        o[i] = out(dec_pow_fx(in(a[i]), Y, yneg, prec, &ok2));
        okv[i] = ok1 && ok2;
    }
}
// y given as a double (exact binary value, like Decimal(length/3))
void t_powd(int n, const TDec* a, const double* y, int prec, TDec* o, int* okv) {
    for (int i = 0; i < n; i++) {
        bool ok2;
        Fx Y = fx_from_double(fabs(y[i]));
        o[i] = out(dec_pow_fx(in(a[i]), Y, y[i] < 0, prec, &ok2));
        okv[i] = ok2;
    }
}
void t_repr(int n, const double* v, TDec* o, int* okv) {
    for (int i = 0; i < n; i++) {
        bool ok;
        o[i] = out(dec_from_double_repr(v[i], &ok));
        okv[i] = ok;
    }
}
// fast 28x28-digit multiply used by the per-codon product; okv = 1 fast path decided, 0 = fell back
void t_hold_fast(int n, const TDec* a, const TDec* b, TDec* o, int* okv) {
    for (int i = 0; i < n; i++) {
        Dec x = in(a[i]), y = in(b[i]);
        HoldFac f;
        holdfac_prepare(y, f);
        U4 c27 = {f.c27[0], f.c27[1], f.c27[2], f.c27[3]}, c28 = {f.c28[0], f.c28[1], f.c28[2], f.c28[3]};
        U4 misc = {f.btop[0], f.btop[1], (u32)f.e, f.ok};
        u32 a0 = x.c.w[0], a1 = x.c.w[1], a2 = x.c.w[2];
        i32 eh = x.e;
        bool ok = f.ok && hold_step_fast(a0, a1, a2, eh, c27, c28, misc);
        Dec r;
        if (ok) {
            r.c.w[0] = a0; r.c.w[1] = a1; r.c.w[2] = a2; r.c.w[3] = 0; r.e = eh; r.neg = 0;
        } else {
            r = dec_mul(x, y);
        }
        okv[i] = ok;
        o[i] = out(r);
    }
}
void t_to_double(int n, const TDec* a, double* o, int* okv) {
    for (int i = 0; i < n; i++) {
        bool ok;
        o[i] = dec_to_double(in(a[i]), &ok);
        okv[i] = ok;
    }
}
// raw fixed-point exp / ln for accuracy measurements: in/out are 7 little-endian u32 limbs (Q32.192)
void t_fx_exp(int n, const u32* t_limbs, const int* neg, u32* out_limbs, int* okv) {
    for (int i = 0; i < n; i++) {
        SFx T;
        for (int j = 0; j < 7; j++) T.m.w[j] = t_limbs[i * 7 + j];
        T.neg = neg[i];
        bool ok;
        Fx v = fx_exp(T, &ok);
        okv[i] = ok;
        for (int j = 0; j < 7; j++) out_limbs[i * 7 + j] = v.w[j];
    }
}
void t_fx_ln(int n, const u32* x_limbs, u32* out_limbs, int* out_neg, int* okv) {
    for (int i = 0; i < n; i++) {
        Fx X;
        for (int j = 0; j < 7; j++) X.w[j] = x_limbs[i * 7 + j];
        bool ok;
        SFx L = fx_ln(X, &ok);
        okv[i] = ok;
        out_neg[i] = L.neg;
        for (int j = 0; j < 7; j++) out_limbs[i * 7 + j] = L.m.w[j];
    }
}
// V = exp(-t); a = round28(V); returns a and ln(a) via the shortcut fx_ln_of_rounded
void t_ln_rounded(int n, const u32* t_limbs, TDec* a_out, u32* ln_limbs, int* ln_neg, int* okv) {
    for (int i = 0; i < n; i++) {
        SFx T;
        for (int j = 0; j < 7; j++) T.m.w[j] = t_limbs[i * 7 + j];
        T.neg = 1;
        bool o1, o2, o3;
        Fx V = fx_exp(T, &o1);
        Dec a = fx_to_dec(V, 28, &o2);
        SFx L = fx_ln_of_rounded(a, V, T, &o3);
        okv[i] = o1 && o2 && o3;
        a_out[i] = out(a);
        ln_neg[i] = L.neg;
        for (int j = 0; j < 7; j++) ln_limbs[i * 7 + j] = L.m.w[j];
    }
}
void t_div_u32(int n, const u32* a, const u32* b, TDec* o) {
    for (int i = 0; i < n; i++) o[i] = out(dec_div_u32(a[i], b[i], ~0ull / b[i], 28));
}
void t_milli(int n, const TDec* a, u32* mag /* n x 8 */, int* okv) {
    for (int i = 0; i < n; i++) {
        Wide<8> m;
        w_zero(m);
        okv[i] = dec_to_milli_int<8>(in(a[i]), m);
        for (int j = 0; j < 8; j++) mag[i * 8 + j] = m.w[j];
    }
}
}
