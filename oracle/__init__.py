"""CPU oracle for the gnina CNN-scoring hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product (gnina_b200/) never does and fails loudly without its CUDA library.
"""
