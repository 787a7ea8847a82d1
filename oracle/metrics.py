"""oracle/metrics.py — TEST INFRASTRUCTURE ONLY: numpy restatement of the two restatable acceptance metrics of
``evaluation_metrics/calculate_intrusive_se_metrics.py`` (reference :37-48 ESTOI via ``pystoi.stoi(extended=True)``,
:90-109 SDR via ``fast_bss_eval.bss_eval_sources(compute_permutation=False, clamp_db=50)``).  Neither ``pystoi`` nor
``fast_bss_eval`` (nor ``pesq`` / ``soxr``) is installed here or installable (no network), so the published algorithms
are restated:

  * ESTOI: Jensen & Taal 2016 as implemented by pystoi 0.4: resample to 10 kHz with the Octave-compatible polyphase
    filter, drop silent frames (40 dB range, 256-sample Hann frames, 50 % overlap), 512-point STFT, 15 one-third octave
    bands from 150 Hz, 30-frame segments, row + column mean/variance normalisation, mean correlation.
  * SDR: BSS-eval v3 SDR with a 512-tap distortion filter, computed the fast_bss_eval way: unit-normalise both signals,
    autocorrelation / cross-correlation by FFT, Toeplitz solve, SDR = 10 log10(coh / (1 - coh)) clamped to +-50 dB.

PESQ (ITU-T P.862 C code inside the ``pesq`` wheel) is NOT restated: **PESQ parity is unpinned** in this container; the
waveform bound (relative L2 <= 1e-3 / 1e-2, i.e. a deviation at least 40-60 dB below the signal) is the operative
guarantee for it.  The acceptance test (north_star: metrics within 0.02 of the reference) evaluates the SAME metric
code on our output and on the reference's output, so small restatement differences cancel in the comparison.
"""
import numpy as np
from scipy.linalg import solve_toeplitz
from scipy.signal import resample_poly

FS = 10000
N_FRAME = 256
NFFT = 512
NUMBAND = 15
MINFREQ = 150
N_SEG = 30
DYN_RANGE = 40
EPS = np.finfo("float").eps


def _resample_window_oct(p, q):
    g = np.gcd(p, q)
    p, q = p // g, q // g
    log10_rejection = -3.0
    stopband_cutoff_f = 1.0 / (2 * max(p, q))
    roll_off_width = stopband_cutoff_f / 10
    rejection_db = -20 * log10_rejection
    L = int(np.ceil((rejection_db - 8) / (28.714 * roll_off_width)))
    t = np.arange(-L, L + 1)
    ideal = 2 * p * stopband_cutoff_f * np.sinc(2 * stopband_cutoff_f * t)
    if 21 <= rejection_db <= 50:
        beta = 0.5842 * (rejection_db - 21) ** 0.4 + 0.07886 * (rejection_db - 21)
    elif rejection_db > 50:
        beta = 0.1102 * (rejection_db - 8.7)
    else:
        beta = 0.0
    return np.kaiser(2 * L + 1, beta) * ideal


def resample_oct(x, p, q):
    h = _resample_window_oct(p, q)
    return resample_poly(x, p, q, window=h / np.sum(h))


def thirdoct(fs, nfft, num_bands, min_freq):
    f = np.linspace(0, fs, nfft + 1)[: nfft // 2 + 1]
    k = np.arange(num_bands, dtype=float)
    cf = 2.0 ** (k / 3.0) * min_freq
    freq_low = min_freq * 2.0 ** ((2 * k - 1) / 6)
    freq_high = min_freq * 2.0 ** ((2 * k + 1) / 6)
    obm = np.zeros((num_bands, len(f)))
    for i in range(num_bands):
        lo = int(np.argmin(np.square(f - freq_low[i])))
        hi = int(np.argmin(np.square(f - freq_high[i])))
        obm[i, lo:hi] = 1
    return obm, cf


def _overlap_and_add(frames, hop):
    n_frames, framelen = frames.shape
    out = np.zeros((n_frames - 1) * hop + framelen)
    for i in range(n_frames):
        out[i * hop: i * hop + framelen] += frames[i]
    return out


def remove_silent_frames(x, y, dyn_range, framelen, hop):
    w = np.hanning(framelen + 2)[1:-1]
    x_frames = np.array([w * x[i:i + framelen] for i in range(0, len(x) - framelen, hop)])
    y_frames = np.array([w * y[i:i + framelen] for i in range(0, len(x) - framelen, hop)])
    x_energies = 20 * np.log10(np.linalg.norm(x_frames, axis=1) + EPS)
    mask = (np.max(x_energies) - dyn_range - x_energies) < 0
    x_frames, y_frames = x_frames[mask], y_frames[mask]
    return _overlap_and_add(x_frames, hop), _overlap_and_add(y_frames, hop)


def _stft(x, win_size, fft_size, overlap=2):
    hop = win_size // overlap
    w = np.hanning(win_size + 2)[1:-1]
    return np.array([np.fft.rfft(w * x[i:i + win_size], n=fft_size) for i in range(0, len(x) - win_size, hop)])


def _row_col_normalize(x, rng):
    xn = x + EPS * rng.standard_normal(x.shape)
    xn = xn - np.mean(xn, axis=-1, keepdims=True)
    xn = xn / np.sqrt(np.sum(np.square(xn), axis=-1, keepdims=True))
    xn = xn + EPS * rng.standard_normal(xn.shape)
    xn = xn - np.mean(xn, axis=1, keepdims=True)
    return xn / np.sqrt(np.sum(np.square(xn), axis=1, keepdims=True))


def estoi(ref, inf, fs):
    """calculate_intrusive_se_metrics.py:37-48 (np.random.seed(0); pystoi.stoi(ref, inf, fs_sig=fs, extended=True))."""
    ref, inf = np.asarray(ref, dtype=np.float64), np.asarray(inf, dtype=np.float64)
    assert ref.shape == inf.shape and ref.ndim == 1
    rng = np.random.RandomState(0)
    if fs != FS:
        ref, inf = resample_oct(ref, FS, fs), resample_oct(inf, FS, fs)
    x, y = remove_silent_frames(ref, inf, DYN_RANGE, N_FRAME, N_FRAME // 2)
    x_spec, y_spec = _stft(x, N_FRAME, NFFT).T, _stft(y, N_FRAME, NFFT).T
    if x_spec.shape[-1] < N_SEG:
        return 1e-5                                             # pystoi: not enough frames -> warning + 1e-5
    obm, _ = thirdoct(FS, NFFT, NUMBAND, MINFREQ)
    x_tob = np.sqrt(obm @ np.square(np.abs(x_spec)))
    y_tob = np.sqrt(obm @ np.square(np.abs(y_spec)))
    xs = np.array([x_tob[:, m - N_SEG:m] for m in range(N_SEG, x_tob.shape[1] + 1)])
    ys = np.array([y_tob[:, m - N_SEG:m] for m in range(N_SEG, x_tob.shape[1] + 1)])
    xn, yn = _row_col_normalize(xs, rng), _row_col_normalize(ys, rng)
    return float(np.sum(xn * yn / N_SEG) / xn.shape[0])


def sdr(ref, inf, filter_length=512, clamp_db=50.0):
    """calculate_intrusive_se_metrics.py:90-109 for one source: BSS-eval v3 SDR (512-tap distortion filter)."""
    ref, inf = np.asarray(ref, dtype=np.float64).reshape(-1), np.asarray(inf, dtype=np.float64).reshape(-1)
    assert ref.shape == inf.shape
    ref = ref / max(np.linalg.norm(ref), 1e-12)
    inf = inf / max(np.linalg.norm(inf), 1e-12)
    n_fft = 2 ** int(np.ceil(np.log2(ref.shape[0] + filter_length)))
    R, E = np.fft.rfft(ref, n=n_fft), np.fft.rfft(inf, n=n_fft)
    acf = np.fft.irfft(R.real ** 2 + R.imag ** 2, n=n_fft)[:filter_length]
    xcorr = np.fft.irfft(R.conj() * E, n=n_fft)[:filter_length]
    sol = solve_toeplitz(acf, xcorr)
    coh = float(np.dot(xcorr, sol))
    lo = 10.0 ** (-clamp_db / 10.0)
    coh = min(max(coh, lo / (1.0 + lo)), 1.0 / (1.0 + lo))     # |SDR| <= clamp_db
    return float(10.0 * np.log10(coh / (1.0 - coh)))
