"""Drop-in mirror of the reference's ``EEG_preprocessing`` package: same module names, function names,
signatures, return order, shapes, dtypes and exceptions -- computed by the CUDA library.

    from eeg2video_b200.EEG_preprocessing.DE_PSD import DE_PSD
    from eeg2video_b200.EEG_preprocessing.segment_raw_signals_200Hz import extract_2s_segment, segment_all_files
    from eeg2video_b200.EEG_preprocessing.segment_sliding_window import seg_sliding_window
    from eeg2video_b200.EEG_preprocessing.extract_DE_PSD_features_1per2s import extract_de_psd_raw
    from eeg2video_b200.EEG_preprocessing.extract_DE_PSD_features_1per1s import extract_de_psd_1s
    from eeg2video_b200.EEG_preprocessing.extract_DE_PSD_features_1per500ms import extract_de_psd_sw

numpy in -> numpy out (one H2D, one kernel, one D2H); torch CUDA tensors in -> torch CUDA tensors out.
"""
