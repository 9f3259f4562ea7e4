#=GENOME_DIFF	1.0
RA	7	.	tiny1	701	0	C	.	allele_frequencies=C:7.59455023e-01,.:2.40544977e-01	fisher_strand_p_value=7.37726e-01	frequency=2.40544977e-01	frequency_lower=1.49266093e-01	frequency_upper=3.50430914e-01	ks_quality_p_value=7.60934e-01	major_base=C	major_cov=21/16	major_frequency=7.59455023e-01	minor_base=.	minor_cov=8/4	new_cov=8/4	ref_cov=21/16	score=12.8	total_cov=29/20
RA	11	.	tiny2	71	1	.	C	allele_frequencies=C:3.33752236e-01,.:6.66247764e-01	fisher_strand_p_value=1.00000e+00	frequency=3.33752236e-01	frequency_lower=1.82385798e-01	frequency_upper=5.12820428e-01	ks_quality_p_value=1.00000e+00	major_base=.	major_cov=12/2	major_frequency=6.66247764e-01	minor_base=C	minor_cov=6/1	new_cov=6/1	ref_cov=12/2	score=10.6	total_cov=18/3
RA	12	.	tiny2	406	0	T	.	allele_frequencies=T:7.14208696e-01,.:2.85791304e-01	fisher_strand_p_value=7.24018e-01	frequency=2.85791304e-01	frequency_lower=1.75680366e-01	frequency_upper=4.15600309e-01	ks_quality_p_value=9.06110e-01	major_base=T	major_cov=15/12	major_frequency=7.14208696e-01	minor_base=.	minor_cov=5/6	new_cov=5/6	ref_cov=15/12	score=12.0	total_cov=20/18
RA	13	.	tiny2	521	0	A	T	allele_frequencies=T:1.00000000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.00000000e+00	frequency_lower=9.70082283e-01	frequency_upper=1.00000000e+00	ks_quality_p_value=1.00000e+00	major_base=T	major_cov=27/18	major_frequency=1.00000000e+00	minor_base=N	minor_cov=0/0	new_cov=27/18	ref_cov=0/0	score=87.9	total_cov=27/18
RA	14	.	tiny2	601	1	.	T	allele_frequencies=T:8.93795261e-01,.:1.06204739e-01	fisher_strand_p_value=6.02597e-01	frequency=8.93795261e-01	frequency_lower=7.88267725e-01	frequency_upper=9.61054805e-01	ks_quality_p_value=5.40922e-01	major_base=T	major_cov=15/17	major_frequency=8.93795261e-01	minor_base=.	minor_cov=3/1	new_cov=15/17	ref_cov=3/1	score=63.9	total_cov=18/18
RA	15	.	tiny2	601	2	.	C	allele_frequencies=C:8.92823609e-01,.:1.07176390e-01	fisher_strand_p_value=6.02597e-01	frequency=8.92823609e-01	frequency_lower=7.87575977e-01	frequency_upper=9.59923330e-01	ks_quality_p_value=7.19175e-01	major_base=C	major_cov=15/17	major_frequency=8.92823609e-01	minor_base=.	minor_cov=3/1	new_cov=15/17	ref_cov=3/1	score=71.6	total_cov=18/18
RA	16	.	tiny2	619	0	G	T	allele_frequencies=T:1.00000000e+00	fisher_strand_p_value=1.00000e+00	frequency=1.00000000e+00	frequency_lower=9.53957557e-01	frequency_upper=1.00000000e+00	ks_quality_p_value=1.00000e+00	major_base=T	major_cov=9/20	major_frequency=1.00000000e+00	minor_base=N	minor_cov=0/0	new_cov=9/20	ref_cov=0/0	score=56.1	total_cov=9/20
RA	17	.	tiny2	715	1	.	A	allele_frequencies=A:4.48689405e-01,.:5.51310590e-01	fisher_strand_p_value=4.87179e-01	frequency=4.48689405e-01	frequency_lower=2.96636442e-01	frequency_upper=6.07945892e-01	ks_quality_p_value=7.40728e-01	major_base=.	major_cov=2/13	major_frequency=5.51310590e-01	minor_base=A	minor_cov=0/12	new_cov=0/12	ref_cov=2/13	score=20.4	total_cov=2/25
MC	2	.	tiny1	1	17	0	8	left_inside_cov=0	left_outside_cov=NA	right_inside_cov=6	right_outside_cov=7
MC	5	.	tiny1	223	625	0	0	left_inside_cov=6	left_outside_cov=8	right_inside_cov=6	right_outside_cov=8
MC	8	.	tiny1	1177	1200	0	0	left_inside_cov=6	left_outside_cov=7	right_inside_cov=0	right_outside_cov=NA
MC	18	.	tiny2	781	800	0	0	left_inside_cov=6	left_outside_cov=7	right_inside_cov=0	right_outside_cov=NA
UN	1	.	tiny1	1	14
UN	3	.	tiny1	16	17
UN	4	.	tiny1	19	19
UN	6	.	tiny1	224	625
UN	9	.	tiny1	1178	1200
UN	10	.	tiny2	1	10
UN	19	.	tiny2	782	800
