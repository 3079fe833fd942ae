"""B200-native (sm_100a) batched inference for the PAUT A-scan signal models of
CSMaus/DefectDetection_viaObjectDetection: drop-in nn.Modules over the libpaut.so C ABI."""
from .modules import (DETECTION, ComplexDetectionModel, DefectDetectionModel, EnhancedSignalSequenceDetector,
                      HybridBinaryModel, ImprovedMultiSignalClassifier, MultiSignalClassifier,
                      MultiSignalClassifier_N, MultiSignalClassifierLegacy, SignalSequenceDetector,
                      TwoStageDefectDetector, load_checkpoint_state, sample_indices)
from .dataio import (json_scan_sequences, json_signal_sets, load_json_volume, load_sequences_pickle,
                     sequence_targets)
from .runtime import (NativeModel, detection_metrics, difference_matrix, gather_windows, get_context, group_nonzero, metrics_confusion,
                      metrics_match, window_table)

__all__ = ["MultiSignalClassifier", "MultiSignalClassifier_N", "DefectDetectionModel", "SignalSequenceDetector",
           "EnhancedSignalSequenceDetector", "TwoStageDefectDetector", "MultiSignalClassifierLegacy",
           "ImprovedMultiSignalClassifier", "HybridBinaryModel", "ComplexDetectionModel", "NativeModel", "get_context",
           "gather_windows", "group_nonzero", "window_table", "json_signal_sets", "json_scan_sequences", "load_json_volume", "load_sequences_pickle", "sequence_targets", "difference_matrix", "detection_metrics", "metrics_match", "metrics_confusion",
           "sample_indices", "load_checkpoint_state", "DETECTION"]
