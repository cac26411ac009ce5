"""Rate control: scalar host logic between block rows (reference encoder/RateControl/*)."""
