"""TF-free / fastavro-free readers and writers for the on-disk contracts either side of the hot path
(SURVEY.md Appendix C): TFRecord Example / SequenceExample in, Photon-ML Avro models and score files out."""
