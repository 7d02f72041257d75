"""HyperSpy wrapper (reference: pguresvt/hspy.py:7-80).  Pure data plumbing around SVT.denoise: a signal
with a 2-D signal space is unfolded to (rows, cols, frames), denoised, and folded back into a copy of the
signal.  HyperSpy itself is not imported here — the wrapper only uses the signal's own methods, so any
object implementing the same duck-typed interface works.
"""
from .svt import SVT


class HSPYSVT(SVT):
    """SVT that accepts HyperSpy signals instead of bare arrays."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._signal_type = None
        self._X = None

    def _prepare_to_denoise(self, signal):
        """Extract the array to denoise, navigation axis last (hspy.py:21-51)."""
        sig_dim = signal.axes_manager.signal_dimension
        if sig_dim == 1:
            self._signal_type = "spectrum"
            self._X = signal._data_aligned_with_axes
        elif sig_dim == 2:
            self._signal_type = "image"
            signal.unfold_navigation_space()
            as_spectrum = signal.as_signal1D(spectral_axis=0)
            self._X = as_spectrum._data_aligned_with_axes
            signal.fold()
        else:
            raise NotImplementedError(f"Expected 1D or 2D signal - got dimension {sig_dim}")

    def denoise(self, signal):
        """Denoise `signal`; returns a new signal titled "Denoised <title>" (hspy.py:53-80)."""
        self._prepare_to_denoise(signal)
        super().denoise(self._X)
        axes = (1, 0) if self._signal_type == "spectrum" else (2, 0, 1)
        out = signal._deepcopy_with_new_data(self.Y_.transpose(axes))
        out.metadata.General.title = f"Denoised {signal.metadata.General.title}".strip()
        return out
