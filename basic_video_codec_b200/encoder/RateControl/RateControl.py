"""Bit budgets and the bits->QP lookup (reference encoder/RateControl/RateControl.py:5-43)."""


def bit_budget_per_frame(ec):
    return ec.targetBR / ec.frame_rate


def calculate_constant_row_bit_budget(remaining_bits, row_idx, ec):
    """RCflag 1: what is left of the frame budget, spread evenly over the rows still to code (:9-20)."""
    rows_left = ec.resolution[1] // ec.block_size - row_idx
    return remaining_bits / rows_left


def calculate_proportional_row_bit_budget(frame, row_idx, ec):
    """RCflag 2/3, second pass: the frame budget split like the first pass spent its bits (:23-30)."""
    first = frame.prev_pass_frame
    if first is None:
        raise ValueError("cant find proportional bit budget as prev_pass_frame not defined")
    share = first.bits_per_row[row_idx] / sum(first.bits_per_row)
    return bit_budget_per_frame(ec) * share, share


def find_rc_qp_for_row(bit_budget, qp_table, frame_type="C", scaling_factor=1):
    """Smallest QP whose expected row size fits the budget, else the largest QP of the table (:34-43)."""
    if frame_type not in ("I", "P", "C"):
        raise ValueError("Invalid frame type. Must be one of 'I', 'P', or 'C'.")
    for qp in sorted(qp_table):
        if qp_table[qp][frame_type] * scaling_factor <= bit_budget:
            return qp
    return max(qp_table)
