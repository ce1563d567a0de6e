"""Readers and writers for pocketkaldi's little-endian model file formats.

These are the formats the reference loaders consume; fixtures written here are
read unchanged by the reference (`pk_load`) and by this repo's C-ABI loader
(`pkb_am_load`). Format sources in the reference:

  VEC0      src/vector.cc:392-425      "VEC0", int32 4*dim+4, int32 dim, payload
  MAT0      src/matrix.cc:287-319      "MAT0", int32 8, int32 rows, int32 cols, rows x VEC0
  LAY0      src/nnet.cc:80-130         "LAY0", int32 4, int32 type [, MAT0 W[out x in], VEC0 b]
  NNT0      src/nnet.cc:132-147        "NNT0", int32 4, int32 num_layers, layers
  tid2pdf   tool/convert_trans.py:20-31  VEC0 with an int32 payload, index 0 unused
  pk::fst_0 src/fst.cc:29-92           32-byte name, int32 size, nstates, narcs, start,
                                       float final[nstates], int32 first_arc[nstates], arcs
  SYM0      src/symbol_table.cc:24-68  "SYM0", int32 size, int32 n, int32 buflen, idx[n], buffer
  .conf     src/configuration.cc:14-71 key = value, '#' comments, paths relative to the file
"""

import os
import struct

import numpy as np

LINEAR, RELU, NORMALIZE, SOFTMAX = 0, 1, 2, 3
MUL = 5  # tool/convert_am.py:16-22; not in the reference reader's enum (src/nnet.h)
SIGMOID = 6  # PKB_LAYER_SIGMOID: not a reference layer type at all (north_star option, parity unpinned)
LAYER_NAMES = {LINEAR: "linear", RELU: "relu", NORMALIZE: "normalize", SOFTMAX: "softmax",
               SIGMOID: "sigmoid"}
FST_SECTION = b"pk::fst_0"


# ----------------------------------------------------------------------------- VEC0 / MAT0
def _vec_bytes(vec, dtype):
    a = np.ascontiguousarray(vec, dtype=dtype)
    return b"VEC0" + struct.pack("<ii", 4 * a.size + 4, a.size) + a.tobytes()


def write_vector(path, vec, dtype="<f4"):
    with open(path, "wb") as fd:
        fd.write(_vec_bytes(vec, dtype))


def _read_exact(fd, n):
    b = fd.read(n)
    if len(b) != n:
        raise IOError("unexpected end of file: %s" % getattr(fd, "name", "?"))
    return b


def _read_vec(fd, dtype="<f4"):
    if _read_exact(fd, 4) != b"VEC0":
        raise ValueError("VEC0 section expected")
    size, dim = struct.unpack("<ii", _read_exact(fd, 8))
    if size != 4 * dim + 4:
        raise ValueError("VEC0: section_size %d != 4*%d+4" % (size, dim))
    return np.frombuffer(_read_exact(fd, 4 * dim), dtype=dtype).copy()


def read_vector(path, dtype="<f4"):
    with open(path, "rb") as fd:
        return _read_vec(fd, dtype)


def _mat_bytes(mat):
    m = np.ascontiguousarray(mat, dtype="<f4")
    out = [b"MAT0", struct.pack("<iii", 8, m.shape[0], m.shape[1])]
    for row in m:
        out.append(_vec_bytes(row, "<f4"))
    return b"".join(out)


def _read_mat(fd):
    if _read_exact(fd, 4) != b"MAT0":
        raise ValueError("MAT0 section expected")
    size, rows, cols = struct.unpack("<iii", _read_exact(fd, 12))
    if size != 8:
        raise ValueError("MAT0: section_size %d != 8" % size)
    m = np.empty((rows, cols), dtype=np.float32)
    for r in range(rows):
        v = _read_vec(fd)
        if v.size != cols:
            raise ValueError("MAT0: row %d has %d cols, %d expected" % (r, v.size, cols))
        m[r] = v
    return m


# ----------------------------------------------------------------------------- NNT0
def write_nnet(path, layers):
    """layers: list of ("linear", W[out x in], b[out]) | ("relu",) | ("normalize",) | ("softmax",)
    | ("mul", v) -- the MUL layer tool/convert_am.py:86-110 writes for a FixedScaleComponent; the
    reference reader rejects it (src/nnet.cc:122-126), libpkb200 folds it into a Linear layer."""
    ids = {v: k for k, v in LAYER_NAMES.items()}
    ids["mul"] = MUL
    with open(path, "wb") as fd:
        fd.write(b"NNT0" + struct.pack("<ii", 4, len(layers)))
        for layer in layers:
            kind = ids[layer[0]]
            fd.write(b"LAY0" + struct.pack("<ii", 4, kind))
            if kind == LINEAR:
                W, b = layer[1], layer[2]
                assert W.shape[0] == b.shape[0]
                fd.write(_mat_bytes(W))
                fd.write(_vec_bytes(b, "<f4"))
            elif kind == MUL:
                fd.write(_vec_bytes(layer[1], "<f4"))


def read_nnet(path):
    layers = []
    with open(path, "rb") as fd:
        if _read_exact(fd, 4) != b"NNT0":
            raise ValueError("NNT0 section expected")
        size, n = struct.unpack("<ii", _read_exact(fd, 8))
        for _ in range(n):
            if _read_exact(fd, 4) != b"LAY0":
                raise ValueError("LAY0 section expected")
            lsize, kind = struct.unpack("<ii", _read_exact(fd, 8))
            if lsize != 4:
                raise ValueError("LAY0: section_size %d != 4" % lsize)
            if kind == LINEAR:
                W = _read_mat(fd)
                b = _read_vec(fd)
                layers.append(("linear", W, b))
            elif kind == MUL:
                layers.append(("mul", _read_vec(fd)))
            elif kind in LAYER_NAMES:
                layers.append((LAYER_NAMES[kind],))
            else:
                raise ValueError("unexpected layer type %d" % kind)
    return layers


# ----------------------------------------------------------------------------- FST / SYM0 / conf
def write_fst(path, num_states, start, finals, arcs):
    """finals: {state: weight}; arcs: list of (src, dst, ilabel, olabel, weight)."""
    arcs = sorted(arcs)
    first = [-1] * num_states
    for i, a in enumerate(arcs):
        if first[a[0]] == -1:
            first[a[0]] = i
    fin = [float("inf")] * num_states
    for s, w in finals.items():
        fin[s] = w
    with open(path, "wb") as fd:
        fd.write(FST_SECTION.ljust(32, b"\0"))
        fd.write(struct.pack("<i", 12 + 8 * num_states + 16 * len(arcs)))
        fd.write(struct.pack("<iii", num_states, len(arcs), start))
        fd.write(struct.pack("<%df" % num_states, *fin))
        fd.write(struct.pack("<%di" % num_states, *first))
        for a in arcs:
            fd.write(struct.pack("<iiif", a[1], a[2], a[3], a[4]))


def write_symbol_table(path, words):
    """words[i] is the string of symbol id i."""
    idx, buf = [], b""
    for w in words:
        idx.append(len(buf))
        buf += w.encode("utf-8") + b"\0"
    with open(path, "wb") as fd:
        fd.write(b"SYM0" + struct.pack("<i", 8 + 4 * len(idx) + len(buf)))
        fd.write(struct.pack("<ii", len(idx), len(buf)))
        fd.write(struct.pack("<%di" % len(idx), *idx))
        fd.write(buf)


def write_conf(path, table):
    with open(path, "w") as fd:
        for k, v in table.items():
            fd.write("%s = %s\n" % (k, v))


def read_conf(path):
    table = {}
    with open(path) as fd:
        for line in fd:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            k, v = line.split("=")
            table[k.strip().lower()] = v.strip()
    return table


def conf_path(conf_file, value):
    if value.startswith("/"):
        return value
    return os.path.join(os.path.dirname(conf_file), value)


# ----------------------------------------------------------------------------- WAV (canonical 44-byte header)
def write_wav16(path, pcm):
    """16 kHz mono 16-bit PCM with the exact 44-byte header src/pcm_reader.cc:67-186 requires."""
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    data = pcm.tobytes()
    with open(path, "wb") as fd:
        fd.write(b"RIFF" + struct.pack("<i", 36 + len(data)) + b"WAVE")
        fd.write(b"fmt " + struct.pack("<ihhiihh", 16, 1, 1, 16000, 32000, 2, 16))
        fd.write(b"data" + struct.pack("<i", len(data)))
        fd.write(data)


def read_wav16(path):
    """Returns int16 samples of a canonical 16 kHz mono 16-bit wav."""
    with open(path, "rb") as fd:
        raw = fd.read()
    if raw[:4] != b"RIFF" or raw[8:12] != b"WAVE" or raw[12:16] != b"fmt ":
        raise ValueError("not a canonical RIFF/WAVE file: %s" % path)
    fmt_size, tag, nch, rate, _, _, bits = struct.unpack("<ihhiihh", raw[16:36])
    if fmt_size != 16 or tag != 1 or nch != 1 or rate != 16000 or bits != 16:
        raise ValueError("16 kHz mono 16-bit PCM expected: %s" % path)
    if raw[36:40] != b"data":
        raise ValueError("data chunk expected at byte 36: %s" % path)
    (n,) = struct.unpack("<i", raw[40:44])
    return np.frombuffer(raw[44:44 + n], dtype="<i2").copy()


# ----------------------------------------------------------------------------- synthetic models
def fold_mul_layers(layers):
    """What libpkb200's loader does with ("mul", v) layers: y = x * v folded into the preceding
    Linear (rows of W and b scaled) when it follows one directly, otherwise into the next Linear
    (columns of W scaled). Returns a layer list the reference reader accepts."""
    out, pending = [], None
    for l in layers:
        if l[0] == "mul":
            v = np.asarray(l[1], np.float32)
            if out and out[-1][0] == "linear":
                W, b = out[-1][1], out[-1][2]
                out[-1] = ("linear", (W * v[:, None]).astype(np.float32), (b * v).astype(np.float32))
            else:
                pending = v if pending is None else (pending * v).astype(np.float32)
        elif l[0] == "linear" and pending is not None:
            out.append(("linear", (l[1] * pending[None, :]).astype(np.float32), l[2]))
            pending = None
        elif pending is not None:
            break
        else:
            out.append(l)
    if pending is not None:
        raise ValueError("mul layer that is not adjacent to a linear layer cannot be folded")
    return out


def make_dnn(rng, in_dim, hidden, num_hidden, num_pdfs, normalize=False, w_scale=None):
    """[Linear, ReLU(, Normalize)] x num_hidden, Linear, Softmax.

    W ~ N(0, 2/fan_in), b ~ N(0, 0.1^2) unless w_scale (uniform +-w_scale) is given.
    """
    layers = []
    d = in_dim
    dims = [hidden] * num_hidden + [num_pdfs]
    for i, o in enumerate(dims):
        if w_scale is None:
            W = (rng.standard_normal((o, d)) * np.sqrt(2.0 / d)).astype(np.float32)
            b = (rng.standard_normal(o) * 0.1).astype(np.float32)
        else:
            W = rng.uniform(-w_scale, w_scale, (o, d)).astype(np.float32)
            b = rng.uniform(-w_scale, w_scale, o).astype(np.float32)
        layers.append(("linear", W, b))
        if i < num_hidden:
            layers.append(("relu",))
            if normalize:
                layers.append(("normalize",))
        d = o
    layers.append(("softmax",))
    return layers


def write_model_dir(out_dir, name, layers, prior, left, right, tid2pdf,
                    cmvn_stats=None, fst=None, words=None):
    """Writes <name>.{nnet,prior,tid2pdf[,fst,sym,cmvn]} and <name>.conf; returns the conf path."""
    os.makedirs(out_dir, exist_ok=True)
    p = lambda ext: os.path.join(out_dir, name + ext)
    write_nnet(p(".nnet"), layers)
    write_vector(p(".prior"), prior)
    write_vector(p(".tid2pdf"), np.asarray(tid2pdf, dtype="<i4"), dtype="<i4")
    table = {
        "nnet": name + ".nnet",
        "prior": name + ".prior",
        "left_context": left,
        "right_context": right,
        "num_pdfs": len(prior),
        "tid2pdf": name + ".tid2pdf",
    }
    if cmvn_stats is not None:
        write_vector(p(".cmvn"), cmvn_stats)
        table["cmvn_stats"] = name + ".cmvn"
    if fst is not None:
        write_fst(p(".fst"), *fst)
        table["fst"] = name + ".fst"
    if words is not None:
        write_symbol_table(p(".sym"), words)
        table["symbol_table"] = name + ".sym"
    write_conf(p(".conf"), table)
    return p(".conf")
