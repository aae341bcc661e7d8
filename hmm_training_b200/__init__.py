"""hmm_training_b200 — B200-native hot path of DemianMArin/HMM_Training.

Reference-shaped modules (same function names and behaviour as the reference's files):
    hmm_training        get_observations, hmm_training, training_with_save (+ batched variants)
    hmm_testing         calculate_log_likelihood, test_hmm (+ score_all)
    hmm_classes         HMMTrained, DataStorageHMM
    codevector_functions  createCodeVector, new_epsilon_centroids, new_adjust_centroids, ...
    codevector_classes  RawDataMFCC, CentroidDataMFCC, DataStorage
Lower level: ``engine`` (array API), ``_lib`` (ctypes binding of libhmmb200.so).
All arithmetic of the hot path runs in hand-written CUDA for sm_100a; there is no CPU fallback.
"""
__version__ = "0.1.0"
