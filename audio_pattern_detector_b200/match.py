"""`match` orchestration, stream adapters and the JSONL contract (reference match.py:24-702).

Only the chunk loop behind ``AudioPatternDetector.find_clip_in_audio`` runs on the GPU;
everything here is thin host code kept API- and output-compatible with the reference so
the CLI and library callers can switch over unchanged.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import struct
import sys
import wave
from pathlib import Path
from typing import Any

import numpy as np

from .audio_clip import AudioClip, AudioStream
from .audio_pattern_detector import AudioPatternDetector, PatternDetectedCallback
from .audio_utils import (DEFAULT_TARGET_SAMPLE_RATE, ffmpeg_get_float32_pcm, pcm_to_float32, resample_audio,
                          seconds_to_time)

_READ_THREADS = 4          # positional-read threads per large file read (_WavFileStreamWrapper._pread)


def _emit_jsonl(event_type: str, **fields: Any) -> None:
    """One JSON object per line, flushed (reference match.py:24-27)."""
    print(json.dumps({"type": event_type, **fields}, ensure_ascii=False), flush=True)


def _read_uint32(stream: Any) -> int:
    data = stream.read(4)
    if len(data) < 4:
        raise ValueError(f"Unexpected EOF reading uint32 (got {len(data)} bytes)")
    return int.from_bytes(data, "little", signed=False)


def _read_patterns_from_multiplexed_stdin(target_sample_rate: int) -> list[AudioClip]:
    """[u32 count] then per pattern [u32 name_len][name][u32 data_len][wav] (reference match.py:38-95)."""
    stdin = sys.stdin.buffer
    count = _read_uint32(stdin)
    if count == 0:
        raise ValueError("No patterns provided in multiplexed stdin")
    if count > 100:
        raise ValueError(f"Too many patterns ({count}), max is 100")
    print(f"Reading {count} pattern(s) from multiplexed stdin...", file=sys.stderr)
    clips: list[AudioClip] = []
    for i in range(count):
        name_len = _read_uint32(stdin)
        if name_len == 0 or name_len > 1024:
            raise ValueError(f"Invalid pattern name length: {name_len}")
        raw_name = stdin.read(name_len)
        if len(raw_name) < name_len:
            raise ValueError(f"Unexpected EOF reading pattern name {i + 1}")
        name = raw_name.decode("utf-8")
        data_len = _read_uint32(stdin)
        if data_len == 0:
            raise ValueError(f"Pattern '{name}' has zero-length data")
        if data_len > 100 * 1024 * 1024:
            raise ValueError(f"Pattern '{name}' data too large: {data_len} bytes")
        blob = stdin.read(data_len)
        if len(blob) < data_len:
            raise ValueError(f"Unexpected EOF reading pattern '{name}' data")
        clip = AudioClip.from_wav_bytes(blob, name, sample_rate=target_sample_rate)
        clips.append(clip)
        print(f"  Loaded pattern '{name}' ({clip.clip_length_seconds():.2f}s)", file=sys.stderr)
    return clips


def _validate_wav_header(stream: Any, target_sample_rate: int) -> tuple[int, int]:
    """Mono PCM16/PCM32/float32 at exactly the target rate (reference match.py:215-283)."""
    tag = stream.read(4)
    if tag != b"RIFF":
        raise ValueError(f"Not a WAV file: expected RIFF, got {tag!r}")
    stream.read(4)
    tag = stream.read(4)
    if tag != b"WAVE":
        raise ValueError(f"Not a WAV file: expected WAVE, got {tag!r}")
    while True:
        cid = stream.read(4)
        if len(cid) < 4:
            raise ValueError("WAV file missing fmt chunk")
        size = struct.unpack("<I", stream.read(4))[0]
        if cid == b"fmt ":
            break
        if len(stream.read(size)) != size:
            raise ValueError("WAV file truncated while skipping chunk")
    fmt = stream.read(size)
    if len(fmt) < 16:
        raise ValueError("WAV fmt chunk too short")
    audio_format, channels, rate, _, _, bits = struct.unpack("<HHIIHH", fmt[:16])
    if audio_format == 1:
        if bits not in (16, 32):
            raise ValueError(f"Expected 16-bit or 32-bit PCM, got {bits}")
    elif audio_format == 3:
        if bits != 32:
            raise ValueError(f"Expected 32-bit float, got {bits}")
    else:
        raise ValueError(f"Expected PCM (1) or IEEE float (3) format, got {audio_format}")
    if channels != 1:
        raise ValueError(f"Expected mono (1 channel), got {channels}")
    if rate != target_sample_rate:
        raise ValueError(f"Expected {target_sample_rate} Hz, got {rate}")
    while True:
        cid = stream.read(4)
        if len(cid) < 4:
            raise ValueError("WAV file missing data chunk")
        raw_size = stream.read(4)
        if len(raw_size) < 4:
            raise ValueError("WAV file truncated")
        if cid == b"data":
            break
        size = struct.unpack("<I", raw_size)[0]
        if len(stream.read(size)) != size:
            raise ValueError("WAV file truncated while skipping chunk")
    return audio_format, bits


class _WavStdinStreamWrapper:
    """float32 byte stream over a WAV arriving on stdin (reference match.py:286-332)."""

    def __init__(self, target_sample_rate: int) -> None:
        fmt, bits = _validate_wav_header(sys.stdin.buffer, target_sample_rate)
        self._dtype = np.dtype(np.float32 if fmt == 3 else (np.int16 if bits == 16 else np.int32))
        label = "float32" if fmt == 3 else f"int{bits}"
        print(f"WAV stdin: {target_sample_rate}Hz, mono, {label}", file=sys.stderr)
        # integer PCM goes to the device as it is and is widened there (AudioPatternDetector._find_clip_in_pcm);
        # float32 WAV data needs no conversion at all and takes the float path
        self.pcm_format = None if fmt == 3 else (self._dtype.itemsize, 1)

    def read_pcm(self, frames: int, /) -> bytes:
        """Raw PCM frames (blocks until ``frames`` frames or end of input, like the reference's read)."""
        return sys.stdin.buffer.read(frames * self._dtype.itemsize)

    def read(self, size: int, /) -> bytes:
        data = sys.stdin.buffer.read((size // 4) * self._dtype.itemsize)
        if not data:
            return b""
        raw = np.frombuffer(data, dtype=self._dtype)
        if self._dtype == np.int16:
            return (raw.astype(np.float32) / np.float32(32768.0)).tobytes()
        if self._dtype == np.int32:
            return (raw.astype(np.float32) / np.float32(2147483648.0)).tobytes()
        return raw.tobytes()


class _WavFileStreamWrapper:
    """float32 byte stream over a WAV file, resampled per read if needed (reference match.py:335-431)."""

    def __init__(self, file_path: str, target_sample_rate: int) -> None:
        self.target_sample_rate = target_sample_rate
        self._validated = False
        self._file_path = file_path
        try:
            self._wav: wave.Wave_read = wave.open(file_path, "rb")
        except (wave.Error, FileNotFoundError, OSError) as e:
            raise ValueError(f"Failed to read WAV file {file_path}: {e}")
        self.input_sample_rate = self._wav.getframerate()
        self._channels = self._wav.getnchannels()
        self._sampwidth = self._wav.getsampwidth()
        self.needs_resample = self.input_sample_rate != target_sample_rate
        self._raw: Any = None
        self._raw_left = 0
        self._raw_pos = 0
        self._readers: Any = None
        # raw 8/16/32-bit integer frames go to the device as they are; the detector widens them there and, when the
        # file's rate differs from the detector's, resamples every chunk read there too (rows N1 / N2,
        # AudioPatternDetector._find_clip_in_pcm)
        self.pcm_format = ((self._sampwidth, self._channels)
                           if self._sampwidth in (1, 2, 4) and self._channels <= 64 else None)
        self.pcm_sample_rate = self.input_sample_rate
        if self._channels != 1:
            print(f"Warning: WAV has {self._channels} channels, will be mixed to mono", file=sys.stderr)

    def _validate_first_chunk(self, audio: np.ndarray) -> None:
        if self._validated or len(audio) == 0:
            return
        self._validated = True
        notes = []
        if np.any(np.isnan(audio)):
            notes.append("Audio contains NaN values - data may be corrupt")
        if np.any(np.isinf(audio)):
            notes.append("Audio contains Inf values - data may be corrupt")
        peak = np.max(np.abs(audio))
        if peak > 1.5:
            notes.append(f"Audio values exceed expected range (max: {peak:.2f})")
        if np.all(audio == 0):
            notes.append("First chunk is all zeros - verify input is correct")
        for n in notes:
            print(f"Warning: {n}", file=sys.stderr)

    def read(self, size: int, /) -> bytes:
        want = size // 4
        frames = int(want * self.input_sample_rate / self.target_sample_rate) if self.needs_resample else want
        raw = self._wav.readframes(frames)
        if not raw:
            return b""
        if self._sampwidth not in (1, 2, 4):
            raise ValueError(f"Unsupported WAV sample width: {self._sampwidth} bytes")
        audio = pcm_to_float32(raw, self._sampwidth, self._channels)
        self._validate_first_chunk(audio)
        if self.needs_resample:
            audio = resample_audio(audio, self.input_sample_rate, self.target_sample_rate)
        return audio.tobytes()

    def readinto_pcm(self, buf: Any, frames: int, /) -> int:
        """Read up to ``frames`` raw PCM frames straight into ``buf`` (a writable byte buffer, e.g. pinned host
        memory) without intermediate bytes objects; returns the frames read.  Uses its own handle on the file's
        ``data`` chunk, found by walking the RIFF chunks."""
        if self._raw is None:
            f = open(self._file_path, "rb")
            head = f.read(12)
            if head[:4] != b"RIFF" or head[8:12] != b"WAVE":
                f.close()
                raise ValueError(f"Not a WAV file: {self._file_path}")
            while True:
                hdr = f.read(8)
                if len(hdr) < 8:
                    f.close()
                    raise ValueError(f"WAV file {self._file_path} has no data chunk")
                size = struct.unpack("<I", hdr[4:])[0]
                if hdr[:4] == b"data":
                    self._raw, self._raw_left = f, min(size, self._wav.getnframes() * self._sampwidth * self._channels)
                    self._raw_pos = f.tell()
                    break
                f.seek(size + (size & 1), 1)
        frame_bytes = self._sampwidth * self._channels
        want = min(frames * frame_bytes, self._raw_left)
        view = memoryview(buf).cast("B")[:want]
        got = self._pread(view)
        got -= got % frame_bytes
        self._raw_pos += got
        self._raw_left -= got
        if got and not self._validated:
            self._validated = True
            if not any(view[:min(got, 1 << 20)]):
                print("Warning: First chunk is all zeros - verify input is correct", file=sys.stderr)
        return got // frame_bytes

    def _pread(self, view: memoryview) -> int:
        """Fill ``view`` from the current position of the data chunk.  Large reads are cut into slices read by a
        few threads with positional reads (``os.preadv`` releases the GIL): one thread copying from the page
        cache into pinned memory runs at ~2 GB/s, well below what the device scans."""
        want = len(view)
        fd, pos = self._raw.fileno(), self._raw_pos
        if want < (8 << 20):
            parts = [(0, want)]
        else:
            k = min(_READ_THREADS, want >> 22)
            step = (want // k + 4095) & ~4095
            parts = [(a, min(a + step, want)) for a in range(0, want, step)]

        def fill(part: tuple[int, int]) -> int:
            a, b = part
            done = a
            while done < b:
                r = os.preadv(fd, [view[done:b]], pos + done)
                if r <= 0:
                    break
                done += r
            return done - a

        if len(parts) == 1:
            return fill(parts[0])
        if self._readers is None:
            from concurrent.futures import ThreadPoolExecutor
            self._readers = ThreadPoolExecutor(_READ_THREADS)
        got = 0
        for (a, b), r in zip(parts, self._readers.map(fill, parts)):
            got += r
            if r < b - a:                                                    # short read: the file ends here
                break
        return got

    def read_pcm(self, frames: int, /) -> bytes:
        """Raw interleaved PCM frames (only meaningful when ``pcm_format`` is set)."""
        raw = self._wav.readframes(frames)
        if raw and not self._validated:
            self._validated = True
            if not any(raw):
                print("Warning: First chunk is all zeros - verify input is correct", file=sys.stderr)
        return raw

    def close(self) -> None:
        self._wav.close()
        if self._readers is not None:
            self._readers.shutdown()
            self._readers = None
        if self._raw is not None:
            self._raw.close()
            self._raw = None


def _detect(stream: Any, name: str, clips: list[AudioClip], sr: int, *, debug_mode: bool,
            on_pattern_detected: PatternDetectedCallback | None, accumulate_results: bool,
            seconds_per_chunk: int | None, debug_dir: str, height_min: float | None, live: bool = False
            ) -> tuple[dict[str, list[float]] | None, float]:
    # a pipe may be a live stream: scan chunk by chunk there, so detections are reported as the reference reports
    # them (after every chunk read); a file is scanned many chunks per device call
    detector = AudioPatternDetector(debug_mode=debug_mode, audio_clips=clips, seconds_per_chunk=seconds_per_chunk,
                                    target_sample_rate=sr, debug_dir=debug_dir, height_min=height_min,
                                    max_batch_chunks=1 if live else None, stream_read_chunks=1 if live else None)
    return detector.find_clip_in_audio(AudioStream(name=name, audio_stream=stream, sample_rate=sr),
                                       on_pattern_detected=on_pattern_detected,
                                       accumulate_results=accumulate_results)


def match_pattern(audio_source: str | None, pattern_files: list[str], debug_mode: bool = False,
                  on_pattern_detected: PatternDetectedCallback | None = None, accumulate_results: bool = True,
                  seconds_per_chunk: int | None = 60, from_stdin: bool = False,
                  target_sample_rate: int | None = None, debug_dir: str = "./tmp",
                  height_min: float | None = None) -> tuple[dict[str, list[float]] | None, float]:
    """Find pattern matches in an audio file or WAV stdin (reference match.py:98-212)."""
    if not from_stdin and (audio_source is None or not os.path.exists(audio_source)):
        raise ValueError(f"Audio {audio_source} does not exist")
    sr = target_sample_rate if target_sample_rate is not None else DEFAULT_TARGET_SAMPLE_RATE
    clips: list[AudioClip] = []
    origin: dict[str, str] = {}
    for path in pattern_files:
        if not os.path.exists(path):
            raise ValueError(f"Pattern {path} does not exist")
        clip = AudioClip.from_audio_file(path, sample_rate=sr)
        if clip.name in origin:
            raise ValueError(f"Duplicate clip name '{clip.name}' from files:\n  - {origin[clip.name]}\n  - {path}\n"
                             "Use --pattern-file with name=path syntax to specify unique names.")
        origin[clip.name] = path
        clips.append(clip)
    if not clips:
        raise ValueError("No pattern clips passed")
    common = dict(debug_mode=debug_mode, on_pattern_detected=on_pattern_detected,
                  accumulate_results=accumulate_results, seconds_per_chunk=seconds_per_chunk, debug_dir=debug_dir)
    if from_stdin:
        wrapper = _WavStdinStreamWrapper(sr)
        print("Finding pattern in audio stream stdin...", file=sys.stderr)
        return _detect(wrapper, "stdin", clips, sr, height_min=height_min, live=True, **common)
    assert audio_source is not None
    name = Path(audio_source).stem
    print(f"Finding pattern in audio file {name}...", file=sys.stderr)
    if audio_source.lower().endswith(".wav"):
        wrapper2 = _WavFileStreamWrapper(audio_source, sr)
        try:
            return _detect(wrapper2, name, clips, sr, height_min=height_min, **common)
        finally:
            wrapper2.close()
    with ffmpeg_get_float32_pcm(audio_source, target_sample_rate=sr, ac=1) as pipe:
        # the reference does not forward height_min on this branch (match.py:199-205); kept as is
        return _detect(pipe, name, clips, sr, height_min=None, live=True, **common)


def _match_pattern_multiplexed_stdin(debug_mode: bool, on_pattern_detected: PatternDetectedCallback | None,
                                     accumulate_results: bool, seconds_per_chunk: int | None,
                                     target_sample_rate: int, debug_dir: str = "./tmp",
                                     height_min: float | None = None
                                     ) -> tuple[dict[str, list[float]] | None, float]:
    """reference match.py:477-521."""
    clips = _read_patterns_from_multiplexed_stdin(target_sample_rate)
    print("Reading WAV audio from stdin...", file=sys.stderr)
    wrapper = _WavStdinStreamWrapper(target_sample_rate)
    return _detect(wrapper, "stdin", clips, target_sample_rate, debug_mode=debug_mode,
                   on_pattern_detected=on_pattern_detected, accumulate_results=accumulate_results,
                   seconds_per_chunk=seconds_per_chunk, debug_dir=debug_dir, height_min=height_min, live=True)


def _make_jsonl_callback(timestamp_format: str = "both") -> PatternDetectedCallback:
    """pattern_detected events; a repeated millisecond for the same clip is dropped (reference match.py:524-551)."""
    last: dict[str, int] = {}

    def callback(clip_name: str, timestamp: float) -> None:
        ms = round(timestamp * 1000)
        if last.get(clip_name) == ms:
            return
        last[clip_name] = ms
        fields: dict[str, Any] = {"clip_name": clip_name}
        if timestamp_format != "formatted":
            fields["timestamp_ms"] = ms
        if timestamp_format != "ms":
            fields["timestamp_formatted"] = seconds_to_time(timestamp)
        _emit_jsonl("pattern_detected", **fields)

    return callback


def _emit_jsonl_end(total_time: float, timestamp_format: str = "both") -> None:
    fields: dict[str, Any] = {}
    if timestamp_format != "formatted":
        fields["total_time_ms"] = round(total_time * 1000)
    if timestamp_format != "ms":
        fields["total_time_formatted"] = seconds_to_time(total_time)
    _emit_jsonl("end", **fields)


def _run_match_with_output(args: argparse.Namespace, pattern_files: list[str], audio_source: str | None,
                           from_stdin: bool = False, seconds_per_chunk: int | None = 60,
                           target_sample_rate: int | None = None, debug_dir: str = "./tmp",
                           height_min: float | None = None) -> tuple[None, float]:
    fmt: str = getattr(args, "timestamp_format", "both")
    _emit_jsonl("start", source="stdin" if from_stdin else (audio_source or "unknown"))
    _, total = match_pattern(audio_source, pattern_files, debug_mode=args.debug,
                             on_pattern_detected=_make_jsonl_callback(fmt), accumulate_results=False,
                             seconds_per_chunk=seconds_per_chunk, from_stdin=from_stdin,
                             target_sample_rate=target_sample_rate, debug_dir=debug_dir, height_min=height_min)
    print(f"Total time processed: {seconds_to_time(seconds=total)}", file=sys.stderr)
    _emit_jsonl_end(total, fmt)
    return None, total


def cmd_match(args: argparse.Namespace) -> None:
    """reference match.py:603-680."""
    text = getattr(args, "chunk_seconds", "60")
    if text.lower() == "auto":
        spc: int | None = None
    else:
        try:
            spc = int(text)
        except ValueError:
            print(f"Error: --chunk-seconds must be 'auto' or a positive integer, got '{text}'", file=sys.stderr)
            sys.exit(1)
    target_sr = getattr(args, "target_sample_rate", None)
    sr = target_sr if target_sr is not None else DEFAULT_TARGET_SAMPLE_RATE
    debug_dir: str = getattr(args, "debug_dir", "./tmp")
    height_min: float | None = getattr(args, "height_min", None)
    fmt: str = getattr(args, "timestamp_format", "both")

    if getattr(args, "multiplexed_stdin", False):
        _emit_jsonl("start", source="multiplexed-stdin")
        _, total = _match_pattern_multiplexed_stdin(debug_mode=args.debug, on_pattern_detected=_make_jsonl_callback(fmt),
                                                    accumulate_results=False, seconds_per_chunk=spc,
                                                    target_sample_rate=sr, debug_dir=debug_dir, height_min=height_min)
        print(f"Total time processed: {seconds_to_time(seconds=total)}", file=sys.stderr)
        _emit_jsonl_end(total, fmt)
        return

    files: list[str] = []
    for folder in args.pattern_folder or []:
        for ext in ("wav", "apd.toml"):
            for f in glob.glob(f"{folder}/*.{ext}"):
                print(f"adding pattern file {f}...", file=sys.stderr)
                files.append(f)
    files.extend(args.pattern_file or [])
    if not files:
        print("Please provide either --pattern-file, --pattern-folder, or --multiplexed-stdin", file=sys.stderr)
        sys.exit(1)
    kw = dict(seconds_per_chunk=spc, target_sample_rate=target_sr, debug_dir=debug_dir, height_min=height_min)
    if args.stdin:
        _run_match_with_output(args, files, None, from_stdin=True, **kw)
    elif args.audio_file:
        _run_match_with_output(args, files, args.audio_file, **kw)
    else:
        print("Please provide an audio file or --stdin or --multiplexed-stdin", file=sys.stderr)
        sys.exit(1)


def cmd_show_config(args: argparse.Namespace) -> None:
    """reference match.py:683-702."""
    target_sr = getattr(args, "target_sample_rate", None)
    if not os.path.exists(args.pattern_file):
        print(f"Error: Pattern {args.pattern_file} does not exist", file=sys.stderr)
        sys.exit(1)
    clips = [AudioClip.from_audio_file(args.pattern_file, sample_rate=target_sr)]
    detector = AudioPatternDetector(audio_clips=clips, debug_mode=False, seconds_per_chunk=None,
                                    target_sample_rate=target_sr)
    print(json.dumps(detector.get_config(), indent=2, ensure_ascii=False))
