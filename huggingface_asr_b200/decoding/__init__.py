"""Mirror of the reference's src/decoding package: only the CTC path (ctc_scorer) is provided."""
