"""CPU oracle of the hot path — TEST INFRASTRUCTURE ONLY (see oracle/floor_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
package; the product package office_person_detection_vit_b200 never does.
"""
