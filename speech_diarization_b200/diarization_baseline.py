"""Drop-in for the OUTPUT FORMAT of /root/reference/diarization_baseline.py (:259-266):
``list[(start, end, speaker)]`` sorted by start and NIST RTTM text.  The pyannote pipeline
itself (segmentation model, HF-gated weights) is out of scope."""
from __future__ import annotations

from pathlib import Path


def segments_to_tuples(segments, label_fmt: str = "SPEAKER_{:02d}") -> list:
    """Segment(start, end, spk) -> (start, end, "SPEAKER_xx") tuples sorted by start, the shape
    diarize_audio returns (diarization_baseline.py:259-261); pyannote names speakers SPEAKER_00…"""
    out = [(float(s.start), float(s.end), label_fmt.format(int(s.spk)) if isinstance(s.spk, (int,)) or
            hasattr(s.spk, "__index__") else s.spk) for s in segments]
    return sorted(out, key=lambda t: (t[0], t[1]))


def rttm_lines(segments: list, uri: str) -> list:
    """One RTTM line per turn, as pyannote.core.Annotation.write_rttm formats it."""
    return [f"SPEAKER {uri} 1 {start:.3f} {end - start:.3f} <NA> <NA> {label} <NA> <NA>\n"
            for start, end, label in segments]


def write_rttm(segments: list, rttm_filepath: str | Path, uri: str | None = None) -> None:
    """diarization_baseline.py:263-265."""
    uri = uri or Path(rttm_filepath).stem
    with open(rttm_filepath, "w") as f:
        f.writelines(rttm_lines(segments, uri))
