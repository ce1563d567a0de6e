"""TEST INFRASTRUCTURE: ctypes bindings for the CPU restatement oracle
(oracle/libpkoracle.so, built from oracle/pk_oracle.c) and for the compiled
reference (oracle/_ref/libpkref.so, built by oracle/build_ref.sh).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this module. The product (pocketkaldi_b200/) never does.
"""

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libpkoracle.so")
REF_SO = os.path.join(HERE, "_ref", "libpkref.so")
REF_CLI = os.path.join(HERE, "_ref", "pocketkaldi_ref")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build():
    """Compiles the restatement (always) and the reference (when its sources exist)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "libpkoracle.so"])
    subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")])


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ----------------------------------------------------------------------------- restatement
class Oracle:
    """The C restatement. All matrices are [T][dim] row-major float32."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        L.pko_num_frames.argtypes = [C.c_int]
        L.pko_hamming.argtypes = [_f32p]
        L.pko_mel_table.argtypes = [_f32p, _i32p, _i32p]
        L.pko_srfft.argtypes = [_f32p, C.c_int]
        L.pko_fbank.argtypes = [_f32p, C.c_int, _f32p]
        L.pko_cmvn.argtypes = [_f32p, C.c_int, _f32p, _f32p]
        L.pko_splice.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
        L.pko_linear.argtypes = [_f32p, C.c_int, C.c_int, _f32p, _f32p, C.c_int, _f32p]
        L.pko_relu.argtypes = [_f32p, C.c_size_t]
        L.pko_normalize.argtypes = [_f32p, C.c_int, C.c_int]
        L.pko_softmax.argtypes = [_f32p, C.c_int, C.c_int]
        L.pko_am_epilogue.argtypes = [_f32p, C.c_int, C.c_int, _f32p]
        L.pko_log_prior.argtypes = [_f32p, C.c_int, _f32p]
        L.pko_scale.argtypes = [_f32p, C.c_size_t, C.c_float]
        L.pko_simple_matmat.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        self.L = L

    def num_frames(self, n):
        return self.L.pko_num_frames(int(n))

    def hamming(self):
        w = np.empty(400, np.float32)
        self.L.pko_hamming(w)
        return w

    def mel_table(self):
        w = np.empty((40, 256), np.float32)
        off = np.empty(40, np.int32)
        wid = np.empty(40, np.int32)
        self.L.pko_mel_table(w, off, wid)
        return w, off, wid

    def srfft(self, x):
        y = _f32(x).copy()
        self.L.pko_srfft(y, y.size)
        return y

    def fbank(self, wave):
        wave = _f32(wave)
        T = self.num_frames(wave.size)
        out = np.empty((T, 40), np.float32)
        if T:
            self.L.pko_fbank(wave, wave.size, out)
        return out

    def cmvn(self, raw, global_stats):
        raw = _f32(raw)
        out = np.empty_like(raw)
        if raw.shape[0]:
            self.L.pko_cmvn(raw, raw.shape[0], _f32(global_stats), out)
        return out

    def splice(self, feats, left, right):
        feats = _f32(feats)
        T, D = feats.shape
        out = np.empty((T, (left + right + 1) * D), np.float32)
        if T:
            self.L.pko_splice(feats, T, D, left, right, out)
        return out

    def linear(self, x, W, b):
        x, W, b = _f32(x), _f32(W), _f32(b)
        y = np.empty((x.shape[0], W.shape[0]), np.float32)
        self.L.pko_linear(x, x.shape[0], x.shape[1], W, b, W.shape[0], y)
        return y

    def nnet(self, x, layers):
        """layers as returned by pocketkaldi_b200.formats.read_nnet (src/nnet.cc:149-163)."""
        x = _f32(x).copy()
        for layer in layers:
            kind = layer[0]
            if kind == "linear":
                x = self.linear(x, layer[1], layer[2])
            elif kind == "relu":
                self.L.pko_relu(x, x.size)
            elif kind == "normalize":
                self.L.pko_normalize(x, x.shape[0], x.shape[1])
            elif kind == "softmax":
                self.L.pko_softmax(x, x.shape[0], x.shape[1])
            else:
                raise ValueError(kind)
        return x

    def log_prior(self, prior):
        prior = _f32(prior)
        out = np.empty_like(prior)
        self.L.pko_log_prior(prior, prior.size, out)
        return out

    def am_compute(self, feats, layers, prior, left, right):
        """AcousticModel::Compute (src/am.cc:90-115): splice -> nnet -> floor/log/-logprior."""
        feats = _f32(feats)
        if feats.shape[0] == 0:
            return np.empty((0, len(prior)), np.float32)
        x = self.splice(feats, left, right)
        p = self.nnet(x, layers)
        self.L.pko_am_epilogue(p, p.shape[0], p.shape[1], self.log_prior(prior))
        return p

    def decodable(self, feats, layers, prior, left, right, prob_scale):
        """pk_decodable_init (src/decodable.cc:8-17): AM compute then scale."""
        p = self.am_compute(feats, layers, prior, left, right)
        self.L.pko_scale(p, p.size, prob_scale)
        return p

    def simple_matmat(self, A, B):
        A, B = _f32(A), _f32(B)
        out = np.empty((A.shape[0], B.shape[1]), np.float32)
        self.L.pko_simple_matmat(A, B, out, A.shape[0], A.shape[1], B.shape[1])
        return out


# ----------------------------------------------------------------------------- compiled reference
def ref_available():
    return os.path.exists(REF_SO)


class Reference:
    """The unmodified reference through oracle/ref_capi.cc."""

    def __init__(self):
        if not ref_available():
            raise RuntimeError("oracle/_ref/libpkref.so missing: run oracle/build_ref.sh where "
                               "/root/reference exists")
        L = C.CDLL(REF_SO)
        L.ref_fbank.argtypes = [_f32p, C.c_int, _f32p, C.c_int]
        L.ref_cmvn.argtypes = [_f32p, C.c_int, _f32p, _f32p]
        L.ref_srfft.argtypes = [_f32p, C.c_int]
        L.ref_gemm.argtypes = [_f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.ref_nnet_propagate.argtypes = [C.c_char_p, _f32p, C.c_int, C.c_int, _f32p, C.c_int,
                                         C.c_char_p, C.c_int]
        L.ref_am_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.ref_am_load.restype = C.c_void_p
        L.ref_am_free.argtypes = [C.c_void_p]
        L.ref_am_num_pdfs.argtypes = [C.c_void_p]
        L.ref_am_tid2pdf.argtypes = [C.c_void_p, C.c_int]
        L.ref_am_compute.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, _f32p]
        L.ref_decodable_fill.argtypes = [C.c_void_p, C.c_float, _f32p, C.c_int, C.c_int, _i32p,
                                         C.c_int, _f32p, _u8p]
        L.ref_decode_wav.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int,
                                     C.POINTER(C.c_float), C.c_char_p, C.c_int]
        L.ref_read_wav.argtypes = [C.c_char_p, _f32p, C.c_int]
        L.ref_time_path.argtypes = [C.c_void_p, _f32p, _i16p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(C.c_longlong), C.POINTER(C.c_double)]
        L.ref_time_path.restype = C.c_double
        self.L = L

    def fbank(self, wave):
        wave = _f32(wave)
        T = self.L.ref_fbank_num_frames(wave.size)
        out = np.empty((T, 40), np.float32)
        got = self.L.ref_fbank(wave, wave.size, out, T)
        assert got == T
        return out

    def cmvn(self, raw, global_stats):
        raw = _f32(raw)
        out = np.empty_like(raw)
        if raw.shape[0]:
            self.L.ref_cmvn(raw, raw.shape[0], _f32(global_stats), out)
        return out

    def srfft(self, x):
        y = _f32(x).copy()
        self.L.ref_srfft(y, y.size)
        return y

    def gemm(self, A, B):
        A, B = _f32(A), _f32(B)
        out = np.zeros((A.shape[0], B.shape[1]), np.float32)
        self.L.ref_gemm(A, B, out, A.shape[0], A.shape[1], B.shape[1])
        return out

    def nnet_propagate(self, nnet_path, x, out_dim):
        x = _f32(x)
        out = np.empty((x.shape[0], out_dim), np.float32)
        err = C.create_string_buffer(512)
        got = self.L.ref_nnet_propagate(nnet_path.encode(), x, x.shape[0], x.shape[1], out,
                                        out.size, err, 512)
        if got < 0:
            raise RuntimeError(err.value.decode())
        assert got == out_dim
        return out

    def am_load(self, conf_path):
        err = C.create_string_buffer(512)
        h = self.L.ref_am_load(conf_path.encode(), err, 512)
        if not h:
            raise RuntimeError(err.value.decode())
        return h

    def am_free(self, h):
        self.L.ref_am_free(h)

    def am_num_pdfs(self, h):
        return self.L.ref_am_num_pdfs(h)

    def am_tid2pdf(self, h, tid):
        return self.L.ref_am_tid2pdf(h, tid)

    def am_compute(self, h, feats):
        feats = _f32(feats)
        P = self.am_num_pdfs(h)
        out = np.empty((feats.shape[0], P), np.float32)
        if feats.shape[0]:
            self.L.ref_am_compute(h, feats, feats.shape[0], feats.shape[1], out)
        return out

    def decodable_fill(self, h, prob_scale, feats, tids):
        feats = _f32(feats)
        tids = np.ascontiguousarray(tids, dtype=np.int32)
        out = np.empty((feats.shape[0], tids.size), np.float32)
        last = np.empty(feats.shape[0], np.uint8)
        self.L.ref_decodable_fill(h, prob_scale, feats, feats.shape[0], feats.shape[1], tids,
                                  tids.size, out, last)
        return out, last

    def decode_wav(self, conf_path, wav_path):
        hyp = C.create_string_buffer(4096)
        err = C.create_string_buffer(512)
        llpf = C.c_float(0)
        rc = self.L.ref_decode_wav(conf_path.encode(), wav_path.encode(), hyp, 4096,
                                   C.byref(llpf), err, 512)
        if rc != 0:
            raise RuntimeError(err.value.decode())
        return hyp.value.decode(), llpf.value

    def read_wav(self, wav_path):
        buf = np.empty(1 << 24, np.float32)
        n = self.L.ref_read_wav(wav_path.encode(), buf, buf.size)
        if n < 0:
            raise RuntimeError("ref_read_wav failed: %s" % wav_path)
        return buf[:n].copy()

    def time_path(self, am_handle, global_stats, pcm_i16, n_threads, repeats=1):
        """pcm_i16: [n_utts][samples]. Returns (seconds, frames, checksum)."""
        pcm = np.ascontiguousarray(pcm_i16, dtype=np.int16)
        frames = C.c_longlong(0)
        chk = C.c_double(0)
        sec = self.L.ref_time_path(am_handle, _f32(global_stats), pcm.reshape(-1), pcm.shape[0],
                                   pcm.shape[1], n_threads, repeats, C.byref(frames),
                                   C.byref(chk))
        return sec, frames.value, chk.value
