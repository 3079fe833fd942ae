"""B200-native (sm_100a) batched inference for the PAUT A-scan signal models of
CSMaus/DefectDetection_viaObjectDetection: drop-in nn.Modules over the libpaut.so C ABI."""
from .modules import (DETECTION, DefectDetectionModel, EnhancedSignalSequenceDetector, MultiSignalClassifier,
                      MultiSignalClassifier_N, SignalSequenceDetector, TwoStageDefectDetector,
                      load_checkpoint_state, sample_indices)
from .runtime import NativeModel, gather_windows, get_context, window_table

__all__ = ["MultiSignalClassifier", "MultiSignalClassifier_N", "DefectDetectionModel", "SignalSequenceDetector",
           "EnhancedSignalSequenceDetector", "TwoStageDefectDetector", "NativeModel", "get_context",
           "gather_windows", "window_table", "sample_indices", "load_checkpoint_state", "DETECTION"]
